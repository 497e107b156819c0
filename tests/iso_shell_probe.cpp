// prints the tables of a2ds::IsoShell for a few sections (host only)
#include <cstdio>

#include "../a2d-shells_b200/host/IsoShell.h"

int main() {
  const double off[3] = {0.0, 0.3, -0.45};
  for (int k = 0; k < 3; k++) {
    a2ds::IsoShell s;
    s.t_offset = off[k];
    s.t = 0.010 + 0.004 * k;
    s.E = 72e9 * (1 + k);
    double Cs[22], eth[9], mom[3];
    s.tangentStiffness(Cs); s.thermalStrain(eth); s.massMoments(mom);
    std::printf("SECTION");
    for (double v : Cs) std::printf(" %.17g", v);
    for (double v : eth) std::printf(" %.17g", v);
    for (double v : mom) std::printf(" %.17g", v);
    std::printf("\n");
  }
  return 0;
}
