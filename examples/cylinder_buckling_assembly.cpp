// The assembly half of examples/cylinder-buckling/mechBuckling.cpp of the reference, on the
// device and without TACS: read the deck (a2ds::MeshLoader = TACSMeshLoader), one isotropic
// shell section for every component (mechBuckling.cpp:41-66), TACSQuad4Shell with the natural
// transform, then the matrices the buckling analysis needs — Kmat and, about a given state,
// the residual and Gmat (src/TACSBuckling.cpp:239-270).  The eigenvalue solve itself stays with
// the reference's solvers (out of scope here): with --dump the BCSR arrays are written for it.
//
//   g++ -std=c++11 -O2 -Iinclude examples/cylinder_buckling_assembly.cpp \
//       -La2d-shells_b200/lib -la2ds_b200 -Wl,-rpath,$PWD/a2d-shells_b200/lib -o cyl_asm
//   ./cyl_asm mech-cylinder.bdf [--dump prefix] [--scale 1e-5]
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../a2d-shells_b200/host/DeviceAssembler.h"
#include "../a2d-shells_b200/host/IsoShell.h"
#include "../a2d-shells_b200/host/MeshLoader.h"

// the seeded state of the benchmark: scale * U(-1, 1) from
// splitmix64(seed ^ (6 id + dof)), keyed on the file's own node number
static double seeded(uint64_t id, int dof, double scale) {
  uint64_t x = ((6 * id + (uint64_t)dof) ^ 12345ull) + 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return scale * (2.0 * ((double)(x >> 11) * (1.0 / 9007199254740992.0)) - 1.0);
}

static void checksum(const char *name, const std::vector<double> &v) {
  double amax = 0.0, wsum = 0.0;
  for (size_t i = 0; i < v.size(); i++) {
    amax = std::fmax(amax, std::fabs(v[i]));
    wsum += v[i] * std::cos((double)(i % 1000003));
  }
  std::printf("CHECK %s n=%zu max=%.17g wsum=%.17g\n", name, v.size(), amax, wsum);
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s deck.bdf [--dump prefix] [--scale s]\n", argv[0]);
    return 2;
  }
  std::string dump;
  double scale = 1e-5;
  for (int i = 2; i + 1 < argc; i += 2) {
    if (!std::strcmp(argv[i], "--dump")) dump = argv[i + 1];
    if (!std::strcmp(argv[i], "--scale")) scale = std::atof(argv[i + 1]);
  }
  try {
    auto t0 = std::chrono::steady_clock::now();
    a2ds::MeshLoader mesh;
    if (mesh.scanBDFFile(argv[1])) {
      std::fprintf(stderr, "scanBDFFile failed: %s\n", mesh.lastError());
      return 1;
    }
    const int nn = mesh.getNumNodes(), ne = mesh.getNumElements(), nc = mesh.getNumComponents();
    auto t1 = std::chrono::steady_clock::now();
    std::printf("deck: %d nodes, %d elements, %d components, %d SPC cards (%.3f s)\n", nn, ne, nc,
                mesh.getNumBCs(), std::chrono::duration<double>(t1 - t0).count());

    a2ds::DeviceAssembler dev(0);
    mesh.loadInto(dev);
    a2ds::IsoShell section;  // the section of the shipped example
    std::vector<double> Cs(22 * nc), eth(9 * nc), mom(3 * nc);
    for (int c = 0; c < nc; c++) {
      section.tangentStiffness(&Cs[22 * c]);
      section.thermalStrain(&eth[9 * c]);
      section.massMoments(&mom[3 * c]);
    }
    dev.setComponents(nc, Cs.data(), eth.data(), nullptr, nullptr, A2DS_TRANSFORM_NATURAL, nullptr);
    dev.setMassMoments(nc, mom.data());

    const int *file_nums = nullptr;
    a2ds_mesh_file_numbers(mesh.handle(), &file_nums, nullptr);
    std::vector<double> u(6 * (size_t)nn), res(6 * (size_t)nn);
    for (int n = 0; n < nn; n++)
      for (int k = 0; k < 6; k++) u[6 * (size_t)n + k] = seeded((uint64_t)file_nums[n], k, scale);
    dev.setVariables(u.data());

    const int kmat = dev.createMat(), gmat = dev.createMat();
    auto t2 = std::chrono::steady_clock::now();
    dev.assembleMatType(A2DS_STIFFNESS_MATRIX, kmat);             // TACSBuckling.cpp:239
    dev.assembleMatType(A2DS_GEOMETRIC_STIFFNESS_MATRIX, gmat);   // :266
    dev.assembleRes(res.data());
    auto t3 = std::chrono::steady_clock::now();
    std::printf("assembled Kmat, Gmat, residual in %.3f ms (incl. first-launch set-up)\n",
                1e3 * std::chrono::duration<double>(t3 - t2).count());
    std::vector<double> K = dev.getValues(kmat), G = dev.getValues(gmat);
    checksum("K", K);
    checksum("G", G);
    checksum("res", res);
    if (!dump.empty()) {
      int nrows = 0;
      a2ds_mat_pattern(dev.context(), kmat, 0, &nrows, nullptr, nullptr);
      std::vector<int> rowp(nrows + 1), cols(K.size() / 36);
      a2ds_mat_pattern(dev.context(), kmat, 0, &nrows, rowp.data(), cols.data());
      FILE *fp = std::fopen((dump + ".bcsr").c_str(), "wb");
      if (!fp) throw std::runtime_error("cannot open the dump file");
      const int64_t head[2] = {nrows, (int64_t)cols.size()};
      std::fwrite(head, sizeof(int64_t), 2, fp);
      std::fwrite(rowp.data(), sizeof(int), rowp.size(), fp);
      std::fwrite(cols.data(), sizeof(int), cols.size(), fp);
      std::fwrite(K.data(), sizeof(double), K.size(), fp);
      std::fwrite(G.data(), sizeof(double), G.size(), fp);
      std::fclose(fp);
      std::printf("wrote %s.bcsr (rowp, cols, K, G)\n", dump.c_str());
    }
    std::printf("CYLINDER_ASSEMBLY_OK\n");
    return 0;
  } catch (const std::exception &e) {
    std::printf("CYLINDER_ASSEMBLY_ERROR %s\n", e.what());
    return 1;
  }
}
