// compiles the C++ mirror of TACSMeshLoader and walks a deck with it (host only)
#include <cstdio>

#include "../a2d-shells_b200/host/MeshLoader.h"

int main(int argc, char **argv) {
  a2ds::MeshLoader ml;
  if (ml.scanBDFFile("/nonexistent/deck.bdf") == 0) return 2;
  printf("FAIL_MESSAGE %s\n", ml.lastError());
  if (argc < 2 || ml.scanBDFFile(argv[1]) != 0) {
    printf("scan failed: %s\n", ml.lastError());
    return 1;
  }
  int nn, ne, nb;
  const int *ptr, *conn, *comp, *bn, *bv, *bp;
  const double *X, *vals;
  ml.getConnectivity(&nn, &ne, &ptr, &conn, &comp, &X);
  ml.getBCs(&nb, &bn, &bv, &bp, &vals);
  long csum = 0;
  for (int k = 0; k < ptr[ne]; k++) csum += (long)conn[k] * (1 + k % 7);
  double xsum = 0.0;
  for (int k = 0; k < 3 * nn; k++) xsum += X[k];
  printf("MESH %d %d %d %d %ld %.17g %d\n", nn, ne, nb, ml.getNumComponents(), csum, xsum, bp[nb]);
  for (int c = 0; c < ml.getNumComponents(); c++)
    printf("COMP %d [%s] [%s]\n", c, ml.getElementDescript(c), ml.getComponentDescript(c));
  const int *nums;
  a2ds_mesh_file_numbers(ml.handle(), &nums, nullptr);
  printf("FIND %d %d %d\n", ml.findNode(nums[0] + 1), ml.findNode(nums[nn - 1] + 1), ml.findNode(-5));
  if (argc > 2) {
    if (ml.writeBinary(argv[2])) return 3;
    a2ds::MeshLoader again;
    if (again.readBinary(argv[2])) return 4;
    printf("BINARY %d %d\n", again.getNumNodes(), again.getNumElements());
  }
  printf("MESH_LOADER_OK\n");
  return 0;
}
