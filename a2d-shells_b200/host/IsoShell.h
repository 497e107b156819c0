// IsoShell.h — host-side constitutive tables of an isotropic shell section, in the form
// a2ds_set_components / a2ds_set_mass_moments take.  Same numbers as the reference's
// TACSIsoShellConstitutive::evalTangentStiffness / evalThermalStrain / evalMassMoments
// (src/constitutive/TACSIsoShellConstitutive.cpp:192-226, 438-456, 120-129) with the plane
// stress stiffness of TACSMaterialProperties::evalTangentStiffness2D
// (src/constitutive/TACSMaterialProperties.cpp:295-303).  In the drop-in flow the shim asks
// the reference's own constitutive object instead; this header is for drivers that have none.
#ifndef A2DS_ISO_SHELL_H
#define A2DS_ISO_SHELL_H

namespace a2ds {

struct IsoShell {
  double E = 72e9, nu = 0.33, rho = 2718.0, cte = 10e-6;  // examples/cylinder-buckling/mechBuckling.cpp:41-46
  double t = 0.010;        // thickness (:58)
  double t_offset = 0.0;   // mid-surface offset as a fraction of the thickness
  double kcorr = 5.0 / 6.0;
  double drill_reg = 10.0; // DRILLING_REGULARIZATION, src/constitutive/TACSShellConstitutive.cpp:59

  // Cs[22] = A[6] B[6] D[6] As[3] drill; symmetric 3x3 stored as [11 12 13 22 23 33]
  void tangentStiffness(double Cs[22]) const {
    const double G = 0.5 * E / (1.0 + nu), Dm = E / (1.0 - nu * nu);
    const double Q[6] = {Dm, nu * Dm, 0.0, Dm, 0.0, G};
    const double inertia = t * t * t / 12.0;
    double *A = &Cs[0], *B = &Cs[6], *D = &Cs[12], *As = &Cs[18];
    for (int i = 0; i < 6; i++) {
      D[i] = inertia * Q[i];
      A[i] = Q[i] * t;
      B[i] = 0.0;
      B[i] += -t_offset * t * A[i];
      D[i] += t_offset * t_offset * t * t * A[i];
    }
    As[0] = kcorr * A[5]; As[1] = 0.0; As[2] = kcorr * A[5];
    Cs[21] = 0.5 * drill_reg * (As[0] + As[2]);
  }
  // thermal strain per unit temperature, 9 strain components
  void thermalStrain(double eth[9]) const {
    for (int i = 0; i < 9; i++) eth[i] = 0.0;
    eth[0] = cte; eth[1] = cte;
  }
  void massMoments(double mom[3]) const {
    mom[0] = rho * t;
    mom[1] = -rho * t * t * t_offset;
    mom[2] = rho * t * t * t * (t_offset * t_offset + 1.0 / 12.0);
  }
};

}  // namespace a2ds
#endif
