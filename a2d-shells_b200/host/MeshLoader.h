// MeshLoader.h — C++ mirror of TACSMeshLoader (src/io/TACSMeshLoader.h:57-148) over the
// a2ds_mesh_* entry points of include/a2ds.h.  Same method names, arguments and return
// conventions as the reference class for the part in front of the assembly path:
// scanBDFFile returns 0 / non-zero ("fail"), the get* calls hand out borrowed pointers that
// live as long as the loader.  createTACS has no counterpart (there is no TACSAssembler on
// the device side): loadInto() puts connectivity, nodes and boundary conditions of a deck of
// 4-node shells into an a2ds::DeviceAssembler instead.  Header only; link with -la2ds_b200.
#ifndef A2DS_MESH_LOADER_H
#define A2DS_MESH_LOADER_H

#include <string>
#include <vector>

#include "DeviceAssembler.h"
#include "a2ds.h"

namespace a2ds {

class MeshLoader {
 public:
  MeshLoader() : mesh_(nullptr) {}
  ~MeshLoader() { a2ds_mesh_free(mesh_); }
  MeshLoader(const MeshLoader &) = delete;
  MeshLoader &operator=(const MeshLoader &) = delete;

  // TACSMeshLoader::scanBDFFile (src/io/TACSMeshLoader.cpp:570); n_threads <= 0: all cores
  int scanBDFFile(const char *file_name, int n_threads = 0) {
    a2ds_mesh_free(mesh_);
    mesh_ = nullptr;
    return a2ds_mesh_read_bdf(file_name, n_threads, &mesh_);
  }
  // the binary container instead of a deck
  int readBinary(const char *file_name) {
    a2ds_mesh_free(mesh_);
    mesh_ = nullptr;
    return a2ds_mesh_read_bin(file_name, &mesh_);
  }
  int writeBinary(const char *file_name) const { return a2ds_mesh_write_bin(mesh_, file_name); }
  const char *lastError() const { return a2ds_last_error(); }

  int getNumComponents() const { return size(5); }
  int getNumNodes() const { return size(0); }
  int getNumElements() const { return size(1); }
  int getNumBCs() const { return size(3); }
  const char *getComponentDescript(int comp_num) const {
    const char *d = nullptr;
    return (mesh_ && !a2ds_mesh_component(mesh_, comp_num, nullptr, &d)) ? d : nullptr;
  }
  const char *getElementDescript(int comp_num) const {
    const char *d = nullptr;
    return (mesh_ && !a2ds_mesh_component(mesh_, comp_num, &d, nullptr)) ? d : nullptr;
  }
  // TACSMeshLoader::getConnectivity / getBCs (:1221-1278): any pointer may be NULL
  void getConnectivity(int *num_nodes, int *num_elements, const int **elem_node_ptr,
                       const int **elem_node_conn, const int **elem_component,
                       const double **Xpts) const {
    if (!mesh_) return;
    a2ds_mesh_sizes(mesh_, num_nodes, num_elements, nullptr, nullptr, nullptr, nullptr);
    a2ds_mesh_connectivity(mesh_, elem_node_ptr, elem_node_conn, elem_component, Xpts);
  }
  void getBCs(int *num_bcs, const int **bc_nodes, const int **bc_vars, const int **bc_ptr,
              const double **bc_vals) const {
    if (!mesh_) return;
    a2ds_mesh_sizes(mesh_, nullptr, nullptr, nullptr, num_bcs, nullptr, nullptr);
    a2ds_mesh_bcs(mesh_, bc_nodes, bc_ptr, bc_vars, bc_vals);
  }
  // file number (1-based, as in the deck) -> node index of the arrays, -1 if absent: the
  // search TACSMeshLoader::getAssemblerNodeNums does (:1200-1215)
  int findNode(int file_node_num) const {
    const int *nums = nullptr;
    if (!mesh_ || a2ds_mesh_file_numbers(mesh_, &nums, nullptr)) return -1;
    int lo = 0, hi = size(0) - 1;
    while (lo <= hi) {
      const int mid = lo + (hi - lo) / 2;
      if (nums[mid] == file_node_num - 1) return mid;
      if (nums[mid] < file_node_num - 1) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
  }

  // connectivity, nodes and boundary conditions of a deck of 4-node shells -> device
  // (the part of TACSMeshLoader::createTACS + TACSCreator that concerns one rank, :1130-1184)
  void loadInto(DeviceAssembler &dev) const {
    const int nn = size(0), ne = size(1), nb = size(3);
    std::vector<int> conn(4 * (size_t)ne), masks(nb);
    std::vector<double> vals(6 * (size_t)nb);
    if (a2ds_mesh_quad4(mesh_, conn.data(), masks.data(), vals.data()))
      throw std::runtime_error(std::string("MeshLoader::loadInto: ") + a2ds_last_error());
    const int *comp = nullptr, *bc_nodes = nullptr;
    const double *X = nullptr;
    a2ds_mesh_connectivity(mesh_, nullptr, nullptr, &comp, &X);
    a2ds_mesh_bcs(mesh_, &bc_nodes, nullptr, nullptr, nullptr);
    dev.setMesh(nn, nn, ne, conn.data(), comp);
    dev.setNodes(X);
    dev.addBCs(nb, bc_nodes, masks.data(), vals.data());
  }
  const a2ds_mesh *handle() const { return mesh_; }

 private:
  int size(int which) const {
    int v[6] = {0, 0, 0, 0, 0, 0};
    if (mesh_) a2ds_mesh_sizes(mesh_, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]);
    return v[which];
  }
  a2ds_mesh *mesh_;
};

}  // namespace a2ds
#endif
