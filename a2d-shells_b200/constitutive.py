"""Host-side constitutive tables shipped to the device once per component.

In the drop-in flow the tables are obtained by calling the reference's own virtuals on
the host (`con->evalTangentStiffness`, `con->evalThermalStrain(theta=1)`), so any
TACSShellConstitutive subclass works.  For the stand-alone benchmark we evaluate the
isotropic shell the same way the reference does.
"""
import numpy as np


def iso_shell_tables(E=72e9, nu=0.33, t=0.010, t_offset=0.0, cte=10e-6, kcorr=5.0 / 6.0,
                     drill_reg=10.0):
    """Cs[22] = [A(6) B(6) D(6) As(3) drill] and unit thermal strain eth[9].

    Follows TACSIsoShellConstitutive::evalTangentStiffness
    (src/constitutive/TACSIsoShellConstitutive.cpp:192-226) and evalThermalStrain
    (:438-456); plane-stress Q from TACSMaterialProperties::evalTangentStiffness2D
    (src/constitutive/TACSMaterialProperties.cpp:295-303): symmetric 3x3 stored as
    [Q11 Q12 Q13 Q22 Q23 Q33]; DRILLING_REGULARIZATION = 10
    (src/constitutive/TACSShellConstitutive.cpp:59)."""
    G = 0.5 * E / (1.0 + nu)
    D_ = E / (1.0 - nu * nu)
    Q = np.array([D_, nu * D_, 0.0, D_, 0.0, G])
    A = np.zeros(6); B = np.zeros(6); D = np.zeros(6)
    inertia = t * t * t / 12.0
    for i in range(6):
        D[i] = inertia * Q[i]
        A[i] = Q[i] * t
        B[i] += -t_offset * t * A[i]
        D[i] += t_offset * t_offset * t * t * A[i]
    As = np.array([kcorr * A[5], 0.0, kcorr * A[5]])
    drill = 0.5 * drill_reg * (As[0] + As[2])
    Cs = np.concatenate([A, B, D, As, [drill]])
    eth = np.zeros(9)
    eth[0] = cte; eth[1] = cte
    return Cs, eth


def iso_mass_moments(rho=2718.0, t=0.010, t_offset=0.0):
    """[m0, m1, m2] of TACSIsoShellConstitutive::evalMassMoments
    (src/constitutive/TACSIsoShellConstitutive.cpp:120-129)."""
    return np.array([rho * t, -rho * t * t * t_offset,
                     rho * t * t * t * (t_offset * t_offset + 1.0 / 12.0)])
