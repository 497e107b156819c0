// partition.cpp — element-wise partition of a global mesh into one rank's sub-mesh and its
// ghost-exchange plan, behind the C ABI of include/a2ds.h (a2ds_partition_*).  Host only.
//
// What the reference spreads over TACSCreator::createTACS (node ownership and numbering,
// src/TACSCreator.cpp:1156-1205: a node belongs to the rank of the first element, in global
// element order, that refers to it; every rank numbers its owned nodes first) and the
// TACSBVecDistribute / TACSBVecIndices pair built in TACSAssembler::initialize (the ghost
// nodes of a rank, sorted by global number, and for each owner the nodes it has to send,
// src/bpmat/TACSBVecDistribute.cpp:280-420).  Here each rank derives its own part straight
// from the global connectivity and the element -> rank array, with no communication: every
// rank sees the same two arrays, so the send list rank a builds for rank b and the receive
// list rank b builds for rank a hold the same global nodes in the same (ascending) order.
//
// Cost: three passes over the 4 n_elems connectivity entries plus sorts of the interface
// sets; 16 M elements take well under a second per rank.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/a2ds.h"

int a2ds_set_error_(const char *msg);  // a2ds.cu

struct a2ds_partition {
  int n_owned = 0;
  std::vector<int> elems;       // global ids of this rank's elements, ascending
  std::vector<int> conn_local;  // 4 per local element, local node indices
  std::vector<int> glob;        // local node -> global node: owned (ascending), then ghosts (ascending)
  std::vector<int> ghost_owner; // owner rank of each ghost
  std::vector<int> peers, send_ptr, send_nodes, recv_ptr, recv_nodes;
  // matrix-halo mode (TACSParallelMat flavour) only:
  bool matrix = false;
  std::vector<int> rowp, cols;   // local pattern: full rows for owned nodes, local couplings for ghosts
  std::vector<int> mat_send_ptr, mat_send_blocks, mat_recv_ptr, mat_recv_blocks;  // per peer
};

static int build_impl(int n_nodes, int n_elems, const int *conn, const int *elem_rank, int n_ranks,
                      int rank, bool matrix, a2ds_partition **out) {
  *out = nullptr;
  if (n_nodes < 0 || n_elems < 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return a2ds_set_error_("a2ds_partition_build: bad sizes");
  for (int e = 0; e < n_elems; e++)
    if (elem_rank[e] < 0 || elem_rank[e] >= n_ranks)
      return a2ds_set_error_("a2ds_partition_build: elem_rank outside [0, n_ranks)");
  for (size_t i = 0; i < 4 * (size_t)n_elems; i++)
    if (conn[i] < 0 || conn[i] >= n_nodes)
      return a2ds_set_error_("a2ds_partition_build: connectivity refers to a node outside [0, n_nodes)");

  // pass 1: owner = rank of the first element that touches the node
  std::vector<int> owner((size_t)n_nodes, -1);
  for (int e = 0; e < n_elems; e++)
    for (int k = 0; k < 4; k++) {
      int &o = owner[conn[4 * (size_t)e + k]];
      if (o < 0) o = elem_rank[e];
    }

  a2ds_partition *p = new a2ds_partition();
  // pass 2: my elements; nodes they use; nodes of mine that other ranks use
  // mark: bit 0 = used by one of my elements; per peer sets are collected as (peer, node) pairs
  std::vector<unsigned char> used((size_t)n_nodes, 0);
  std::vector<uint64_t> wanted;  // (peer << 32 | global node) for my nodes used by `peer`
  p->matrix = matrix;
  for (int e = 0; e < n_elems; e++) {
    const int r = elem_rank[e];
    const int *en = &conn[4 * (size_t)e];
    if (r == rank) {
      p->elems.push_back(e);
      for (int k = 0; k < 4; k++) used[en[k]] = 1;
    } else {
      for (int k = 0; k < 4; k++)
        if (owner[en[k]] == rank) wanted.push_back(((uint64_t)r << 32) | (uint32_t)en[k]);
    }
    if (matrix) {
      // The owner of a node holds the node's whole matrix row: every node that shares an
      // element (of any rank) with one of MY nodes is a column of mine, hence local here
      // (the reference's external column map); and my nodes are ghosts wherever a node of
      // another owner shares an element with them.
      bool mine = false;
      for (int k = 0; k < 4; k++) mine = mine || owner[en[k]] == rank;
      if (mine)
        for (int k = 0; k < 4; k++) {
          used[en[k]] = 1;
          const int o = owner[en[k]];
          if (o != rank)
            for (int j = 0; j < 4; j++)
              if (owner[en[j]] == rank) wanted.push_back(((uint64_t)o << 32) | (uint32_t)en[j]);
        }
    }
  }
  std::sort(wanted.begin(), wanted.end());
  wanted.erase(std::unique(wanted.begin(), wanted.end()), wanted.end());

  // local numbering: owned nodes (ascending global number), then ghosts (ascending)
  std::vector<int> local_of((size_t)n_nodes, -1);
  for (int g = 0; g < n_nodes; g++)
    if (owner[g] == rank) {
      local_of[g] = (int)p->glob.size();
      p->glob.push_back(g);
    }
  p->n_owned = (int)p->glob.size();
  for (int g = 0; g < n_nodes; g++)
    if (used[g] && owner[g] != rank) {
      local_of[g] = (int)p->glob.size();
      p->glob.push_back(g);
      p->ghost_owner.push_back(owner[g]);
    }
  p->conn_local.resize(4 * p->elems.size());
  for (size_t i = 0; i < p->elems.size(); i++)
    for (int k = 0; k < 4; k++)
      p->conn_local[4 * i + k] = local_of[conn[4 * (size_t)p->elems[i] + k]];

  // peers: ranks that own one of my ghosts or use one of my nodes
  std::vector<char> is_peer((size_t)n_ranks, 0);
  for (int o : p->ghost_owner) is_peer[o] = 1;
  for (uint64_t w : wanted) is_peer[(size_t)(w >> 32)] = 1;
  for (int r = 0; r < n_ranks; r++)
    if (is_peer[r]) p->peers.push_back(r);
  p->send_ptr.assign(1, 0);
  p->recv_ptr.assign(1, 0);
  size_t wi = 0;
  for (int r : p->peers) {
    // nodes of mine that rank r reads as ghosts (ascending global number)
    while (wi < wanted.size() && (int)(wanted[wi] >> 32) < r) wi++;
    for (; wi < wanted.size() && (int)(wanted[wi] >> 32) == r; wi++)
      p->send_nodes.push_back(local_of[(uint32_t)wanted[wi]]);
    p->send_ptr.push_back((int)p->send_nodes.size());
    // my ghosts owned by rank r (the ghost block is already ascending)
    for (size_t k = 0; k < p->ghost_owner.size(); k++)
      if (p->ghost_owner[k] == r) p->recv_nodes.push_back(p->n_owned + (int)k);
    p->recv_ptr.push_back((int)p->recv_nodes.size());
  }
  if (matrix) {
    // ---- local pattern: rows of my elements' nodes (all of them), and the full row of every
    //      owned node (contributions of any rank's elements) — count, fill, sort + unique
    const int nl = (int)p->glob.size();
    std::vector<int> ptr(nl + 1, 0);
    auto for_rows = [&](auto &&visit) {
      for (int e = 0; e < n_elems; e++) {
        const int *en = &conn[4 * (size_t)e];
        const bool my_elem = elem_rank[e] == rank;
        for (int k = 0; k < 4; k++)
          if (my_elem || owner[en[k]] == rank) visit(local_of[en[k]], en);
      }
    };
    for_rows([&](int row, const int *) { ptr[row + 1] += 4; });
    for (int i = 0; i < nl; i++) ptr[i + 1] += ptr[i];
    std::vector<int> tmp(ptr[nl]), fill(ptr.begin(), ptr.end() - 1);
    for_rows([&](int row, const int *en) {
      for (int j = 0; j < 4; j++) tmp[fill[row]++] = local_of[en[j]];
    });
    p->rowp.assign(nl + 1, 0);
    for (int r = 0; r < nl; r++) {
      std::sort(tmp.begin() + ptr[r], tmp.begin() + ptr[r + 1]);
      const int nu = (int)(std::unique(tmp.begin() + ptr[r], tmp.begin() + ptr[r + 1]) - (tmp.begin() + ptr[r]));
      p->rowp[r + 1] = p->rowp[r] + nu;
    }
    p->cols.resize(p->rowp[nl]);
    for (int r = 0; r < nl; r++)
      std::copy(tmp.begin() + ptr[r], tmp.begin() + ptr[r] + (p->rowp[r + 1] - p->rowp[r]),
                p->cols.begin() + p->rowp[r]);
    auto block_of = [&](int grow, int gcol) {
      const int lr = local_of[grow], lc = local_of[gcol];
      const int *b = p->cols.data() + p->rowp[lr], *e = p->cols.data() + p->rowp[lr + 1];
      const int *it = std::lower_bound(b, e, lc);
      return (it != e && *it == lc) ? (int)(it - p->cols.data()) : -1;
    };
    // ---- blocks that travel: (peer, global row, global column), unique, in that order on
    //      both sides.  I send the blocks my elements add to rows another rank owns; I receive
    //      the blocks other ranks' elements add to rows I own.
    struct Key { int peer, row, col; };
    auto less = [](const Key &a, const Key &b) {
      return a.peer != b.peer ? a.peer < b.peer : (a.row != b.row ? a.row < b.row : a.col < b.col);
    };
    auto same = [](const Key &a, const Key &b) { return a.peer == b.peer && a.row == b.row && a.col == b.col; };
    std::vector<Key> snd, rcv;
    for (int e = 0; e < n_elems; e++) {
      const int *en = &conn[4 * (size_t)e];
      const int r = elem_rank[e];
      for (int k = 0; k < 4; k++) {
        const int o = owner[en[k]];
        if (r == rank && o != rank)
          for (int j = 0; j < 4; j++) snd.push_back({o, en[k], en[j]});
        if (r != rank && o == rank)
          for (int j = 0; j < 4; j++) rcv.push_back({r, en[k], en[j]});
      }
    }
    for (auto *v : {&snd, &rcv}) {
      std::sort(v->begin(), v->end(), less);
      v->erase(std::unique(v->begin(), v->end(), same), v->end());
    }
    p->mat_send_ptr.assign(1, 0);
    p->mat_recv_ptr.assign(1, 0);
    size_t si = 0, ri = 0;
    for (int r : p->peers) {
      for (; si < snd.size() && snd[si].peer <= r; si++)
        if (snd[si].peer == r) p->mat_send_blocks.push_back(block_of(snd[si].row, snd[si].col));
      p->mat_send_ptr.push_back((int)p->mat_send_blocks.size());
      for (; ri < rcv.size() && rcv[ri].peer <= r; ri++)
        if (rcv[ri].peer == r) p->mat_recv_blocks.push_back(block_of(rcv[ri].row, rcv[ri].col));
      p->mat_recv_ptr.push_back((int)p->mat_recv_blocks.size());
    }
    for (int b : p->mat_send_blocks)
      if (b < 0) { delete p; return a2ds_set_error_("a2ds_partition_build: internal: a block to send is not in the pattern"); }
    for (int b : p->mat_recv_blocks)
      if (b < 0) { delete p; return a2ds_set_error_("a2ds_partition_build: internal: an arriving block is not in the pattern"); }
  }
  *out = p;
  return 0;
}

// Element -> rank by recursive coordinate bisection of the element centroids: the set is split
// at the median along its longest axis into two parts sized in proportion to the ranks each
// side gets, recursively.  Deterministic, balanced to within one element per split, compact
// parts (short interfaces) on shell structures.  Stands in for the METIS call of
// TACSCreator::partitionMesh (src/TACSCreator.cpp:923, METIS calls :1118-1125) — the reference's partitioner is an
// external library; any element -> rank array works with a2ds_partition_build.
namespace {
void rcb(std::vector<int> &ids, int lo, int hi, int r0, int r1, const double *cen, int *elem_rank) {
  if (r1 - r0 == 1 || hi - lo <= 0) {
    for (int i = lo; i < hi; i++) elem_rank[ids[i]] = r0;
    return;
  }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = lo; i < hi; i++)
    for (int k = 0; k < 3; k++) {
      mn[k] = std::min(mn[k], cen[3 * (size_t)ids[i] + k]);
      mx[k] = std::max(mx[k], cen[3 * (size_t)ids[i] + k]);
    }
  int ax = 0;
  for (int k = 1; k < 3; k++)
    if (mx[k] - mn[k] > mx[ax] - mn[ax]) ax = k;
  const int rm = r0 + (r1 - r0) / 2;
  const int mid = lo + (int)((long long)(hi - lo) * (rm - r0) / (r1 - r0));
  // ties broken by element number: the split does not depend on the input order of `ids`
  std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int a, int b) {
    const double ca = cen[3 * (size_t)a + ax], cb = cen[3 * (size_t)b + ax];
    return ca < cb || (ca == cb && a < b);
  });
  rcb(ids, lo, mid, r0, rm, cen, elem_rank);
  rcb(ids, mid, hi, rm, r1, cen, elem_rank);
}
}  // namespace

static int rcb_impl(int n_nodes, int n_elems, const int *conn, const double *X, int n_ranks,
                    int *elem_rank) {
  if (n_nodes < 0 || n_elems < 0 || n_ranks < 1) return a2ds_set_error_("a2ds_partition_rcb: bad sizes");
  std::vector<double> cen(3 * (size_t)n_elems);
  for (int e = 0; e < n_elems; e++)
    for (int k = 0; k < 3; k++) {
      double c = 0.0;
      for (int i = 0; i < 4; i++) {
        const int n = conn[4 * (size_t)e + i];
        if (n < 0 || n >= n_nodes)
          return a2ds_set_error_("a2ds_partition_rcb: connectivity refers to a node outside [0, n_nodes)");
        c += X[3 * (size_t)n + k];
      }
      cen[3 * (size_t)e + k] = 0.25 * c;
    }
  std::vector<int> ids(n_elems);
  for (int e = 0; e < n_elems; e++) ids[e] = e;
  rcb(ids, 0, n_elems, 0, n_ranks, cen.data(), elem_rank);
  return 0;
}

extern "C" void a2ds_partition_free(a2ds_partition *p) { delete p; }

extern "C" int a2ds_partition_sizes(const a2ds_partition *p, int *n_local_nodes, int *n_owned,
                                    int *n_local_elems, int *n_peers, int *n_send, int *n_recv) {
  if (!p) return a2ds_set_error_("a2ds_partition_sizes: null partition");
  if (n_local_nodes) *n_local_nodes = (int)p->glob.size();
  if (n_owned) *n_owned = p->n_owned;
  if (n_local_elems) *n_local_elems = (int)p->elems.size();
  if (n_peers) *n_peers = (int)p->peers.size();
  if (n_send) *n_send = (int)p->send_nodes.size();
  if (n_recv) *n_recv = (int)p->recv_nodes.size();
  return 0;
}

extern "C" int a2ds_partition_mesh(const a2ds_partition *p, const int **elems,
                                   const int **conn_local, const int **glob,
                                   const int **ghost_owner) {
  if (!p) return a2ds_set_error_("a2ds_partition_mesh: null partition");
  if (elems) *elems = p->elems.data();
  if (conn_local) *conn_local = p->conn_local.data();
  if (glob) *glob = p->glob.data();
  if (ghost_owner) *ghost_owner = p->ghost_owner.data();
  return 0;
}

extern "C" int a2ds_partition_halo(const a2ds_partition *p, const int **peers,
                                   const int **send_ptr, const int **send_nodes,
                                   const int **recv_ptr, const int **recv_nodes) {
  if (!p) return a2ds_set_error_("a2ds_partition_halo: null partition");
  if (peers) *peers = p->peers.data();
  if (send_ptr) *send_ptr = p->send_ptr.data();
  if (send_nodes) *send_nodes = p->send_nodes.data();
  if (recv_ptr) *recv_ptr = p->recv_ptr.data();
  if (recv_nodes) *recv_nodes = p->recv_nodes.data();
  return 0;
}

// mesh and halo of this rank into a device context: a2ds_set_mesh (with the components of the
// local elements picked from the global elem_comp array, may be NULL) + a2ds_set_halo
extern "C" int a2ds_partition_apply(a2ds_ctx *ctx, const a2ds_partition *p, const int *elem_comp) {
  if (!p) return a2ds_set_error_("a2ds_partition_apply: null partition");
  std::vector<int> comp;
  if (elem_comp) {
    comp.resize(p->elems.size());
    for (size_t i = 0; i < p->elems.size(); i++) comp[i] = elem_comp[p->elems[i]];
  }
  if (a2ds_set_mesh(ctx, (int)p->glob.size(), p->n_owned, (int)p->elems.size(),
                    p->conn_local.data(), elem_comp ? comp.data() : nullptr))
    return 1;
  return a2ds_set_halo(ctx, (int)p->peers.size(), p->peers.data(), p->send_ptr.data(),
                       p->send_nodes.data(), p->recv_ptr.data(), p->recv_nodes.data());
}

// no exception leaves the C ABI: out-of-memory and the like become an error return
extern "C" int a2ds_partition_build(int n_nodes, int n_elems, const int *conn,
                                    const int *elem_rank, int n_ranks, int rank,
                                    a2ds_partition **out) {
  try {
    return build_impl(n_nodes, n_elems, conn, elem_rank, n_ranks, rank, false, out);
  } catch (const std::exception &e) {
    return a2ds_set_error_((std::string("a2ds_partition_build: ") + e.what()).c_str());
  }
}
extern "C" int a2ds_partition_build_matrix(int n_nodes, int n_elems, const int *conn,
                                           const int *elem_rank, int n_ranks, int rank,
                                           a2ds_partition **out) {
  try {
    return build_impl(n_nodes, n_elems, conn, elem_rank, n_ranks, rank, true, out);
  } catch (const std::exception &e) {
    return a2ds_set_error_((std::string("a2ds_partition_build_matrix: ") + e.what()).c_str());
  }
}

extern "C" int a2ds_partition_matrix(const a2ds_partition *p, const int **rowp, const int **cols,
                                     const int **send_ptr, const int **send_blocks,
                                     const int **recv_ptr, const int **recv_blocks) {
  if (!p || !p->matrix)
    return a2ds_set_error_("a2ds_partition_matrix: the partition was not built with a2ds_partition_build_matrix");
  if (rowp) *rowp = p->rowp.data();
  if (cols) *cols = p->cols.data();
  if (send_ptr) *send_ptr = p->mat_send_ptr.data();
  if (send_blocks) *send_blocks = p->mat_send_blocks.data();
  if (recv_ptr) *recv_ptr = p->mat_recv_ptr.data();
  if (recv_blocks) *recv_blocks = p->mat_recv_blocks.data();
  return 0;
}

// a matrix over the local nodes with this pattern whose ghost-row blocks travel to their
// owners inside every assemble call: a2ds_mat_create + a2ds_mat_set_halo
extern "C" int a2ds_partition_create_mat(a2ds_ctx *ctx, const a2ds_partition *p, int *mat) {
  if (!p || !p->matrix)
    return a2ds_set_error_("a2ds_partition_create_mat: the partition was not built with a2ds_partition_build_matrix");
  const int nl = (int)p->glob.size();
  std::vector<int> ident(nl);
  for (int i = 0; i < nl; i++) ident[i] = i;
  const int *rp = p->rowp.data(), *cl = p->cols.data(), *id = ident.data();
  const int one = 1;
  if (a2ds_mat_create(ctx, 1, &nl, &rp, &cl, &id, &id, &one, mat)) return 1;
  return a2ds_mat_set_halo(ctx, *mat, (int)p->peers.size(), p->peers.data(), p->mat_send_ptr.data(),
                           p->mat_send_blocks.data(), p->mat_recv_ptr.data(),
                           p->mat_recv_blocks.data());
}
extern "C" int a2ds_partition_rcb(int n_nodes, int n_elems, const int *conn, const double *X,
                                  int n_ranks, int *elem_rank) {
  try {
    return rcb_impl(n_nodes, n_elems, conn, X, n_ranks, elem_rank);
  } catch (const std::exception &e) {
    return a2ds_set_error_((std::string("a2ds_partition_rcb: ") + e.what()).c_str());
  }
}

// ---- halo plan from the reference's own exchange plan --------------------------------------
// TACSBVecDistribute (src/bpmat/TACSBVecDistribute.h:156-172) describes the ghost exchange of a
// rank with two slab lists of GLOBAL node numbers: ext_vars (sorted; slab i = the nodes this
// rank reads from their owner ext_proc[i]) and req_vars (slab i = the owned nodes rank
// req_proc[i] reads from this rank).  In the local numbering of the device path (owned nodes
// first, g - lo; then the ghosts in the order of ext_vars, exactly the order of TACSBVec's
// external array, src/bpmat/TACSBVec.cpp:982-1039) this becomes the peer / send / receive
// lists of a2ds_set_halo.  Pure host arithmetic: what the drop-in binding calls once per
// assembler on a multi-rank run (host/tacs_shim.cpp).
static int halo_from_distribute_impl(int lo, int n_owned, int n_ext_proc, const int *ext_proc,
                                     const int *ext_ptr, const int *ext_count, int n_req_proc,
                                     const int *req_proc, const int *req_ptr, const int *req_count,
                                     const int *req_vars, int *n_peers, int *peer_rank,
                                     int *send_ptr, int *send_nodes, int *recv_ptr, int *recv_nodes) {
  if (n_ext_proc < 0 || n_req_proc < 0 || n_owned < 0)
    return a2ds_set_error_("a2ds_halo_from_distribute: bad sizes");
  std::vector<int> peers;
  for (int i = 0; i < n_ext_proc; i++) peers.push_back(ext_proc[i]);
  for (int i = 0; i < n_req_proc; i++) peers.push_back(req_proc[i]);
  std::sort(peers.begin(), peers.end());
  peers.erase(std::unique(peers.begin(), peers.end()), peers.end());
  int ns = 0, nr = 0;
  for (size_t k = 0; k < peers.size(); k++) {
    peer_rank[k] = peers[k];
    send_ptr[k] = ns; recv_ptr[k] = nr;
    for (int i = 0; i < n_req_proc; i++)
      if (req_proc[i] == peers[k])
        for (int j = 0; j < req_count[i]; j++) {
          const int g = req_vars[req_ptr[i] + j];
          if (g < lo || g >= lo + n_owned)
            return a2ds_set_error_("a2ds_halo_from_distribute: a requested node is not owned by this rank");
          send_nodes[ns++] = g - lo;
        }
    for (int i = 0; i < n_ext_proc; i++)
      if (ext_proc[i] == peers[k])
        for (int j = 0; j < ext_count[i]; j++) recv_nodes[nr++] = n_owned + ext_ptr[i] + j;
  }
  send_ptr[peers.size()] = ns; recv_ptr[peers.size()] = nr;
  *n_peers = (int)peers.size();
  return 0;
}
extern "C" int a2ds_halo_from_distribute(int lo, int n_owned, int n_ext_proc, const int *ext_proc,
                                         const int *ext_ptr, const int *ext_count, int n_req_proc,
                                         const int *req_proc, const int *req_ptr,
                                         const int *req_count, const int *req_vars, int *n_peers,
                                         int *peer_rank, int *send_ptr, int *send_nodes,
                                         int *recv_ptr, int *recv_nodes) {
  try {
    return halo_from_distribute_impl(lo, n_owned, n_ext_proc, ext_proc, ext_ptr, ext_count, n_req_proc,
                                     req_proc, req_ptr, req_count, req_vars, n_peers, peer_rank,
                                     send_ptr, send_nodes, recv_ptr, recv_nodes);
  } catch (const std::exception &e) {
    return a2ds_set_error_((std::string("a2ds_halo_from_distribute: ") + e.what()).c_str());
  }
}
