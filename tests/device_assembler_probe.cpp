// compiles and (on a GPU box) runs the C++ side-car on a 2x2-element plate
#include <cmath>
#include <cstdio>
#include <vector>

#include "../a2d-shells_b200/host/DeviceAssembler.h"

int main() {
  try {
    a2ds::DeviceAssembler dev(0);
    const int nx = 2, nn = 9, ne = 4;
    std::vector<int> conn;
    for (int j = 0; j < nx; j++)
      for (int i = 0; i < nx; i++) {
        int n0 = j * (nx + 1) + i;
        conn.insert(conn.end(), {n0, n0 + 1, n0 + nx + 1, n0 + nx + 2});
      }
    std::vector<double> X(3 * nn), u(6 * nn);
    for (int n = 0; n < nn; n++) {
      X[3 * n] = 0.1 * (n % 3); X[3 * n + 1] = 0.1 * (n / 3); X[3 * n + 2] = 0.01 * std::sin(1.0 * n);
      for (int k = 0; k < 6; k++) u[6 * n + k] = 1e-5 * std::cos(1.0 + 6 * n + k);
    }
    double Cs[22] = {0}, eth[9] = {0};
    Cs[0] = Cs[3] = 8e8; Cs[1] = 2.6e8; Cs[5] = 2.7e8; Cs[12] = Cs[15] = 6.7e3; Cs[13] = 2.2e3;
    Cs[17] = 2.25e3; Cs[18] = Cs[20] = 2.25e8; Cs[21] = 2.25e9;
    dev.setMesh(nn, nn, ne, conn.data(), nullptr);
    dev.setNodes(X.data());
    dev.setComponents(1, Cs, eth, nullptr, nullptr, A2DS_TRANSFORM_NATURAL, nullptr);
    dev.setVariables(u.data());
    int k = dev.createMat(), g = dev.createMat();
    std::vector<double> r(6 * nn);
    dev.assembleJacobian(1.0, 0.0, 0.0, r.data(), k);
    dev.assembleMatType(A2DS_GEOMETRIC_STIFFNESS_MATRIX, g);
    std::vector<double> K = dev.getValues(k), y(6 * nn);
    dev.mult(k, 0, nn, u.data(), y.data());
    double err = 0, nrm = 0;
    for (int i = 0; i < 6 * nn; i++) { err = std::fmax(err, std::fabs(y[i] - r[i])); nrm = std::fmax(nrm, std::fabs(r[i])); }
    std::printf("DEVICE_ASSEMBLER_OK K u = r to %.2e (|r| %.3e, %zu values)\n", err / nrm, nrm, K.size());
    return err / nrm < 1e-12 ? 0 : 2;
  } catch (const std::exception &e) {
    std::printf("DEVICE_ASSEMBLER_ERROR %s\n", e.what());
    return 1;
  }
}
