// Batch gather from a device-resident dataset (the collate step of the DataLoader, gnn_train.py:387-394).
//
// The whole dataset sits in HBM as concatenated per-sample arrays; a batch is B sample ranges of each.  ONE launch copies
// the node rows (float64 coordinates, mean stress, local stress, labels), the face columns (graph-local node ids stay
// local) and the divergence-operator triplets (rows re-based to the batch's node numbering) into contiguous batch
// arrays, and writes the PyG `batch` vector.  Output = exactly what batcher.batch_from_host builds from host arrays.
#include "pdg_common.cuh"

namespace pdg {

struct ResidentGatherArgs {
  // dataset (device)
  const double* pos64;        // [Ntot][2]
  const float* mean_stress;   // [Ntot][3]
  const float* local_stress;  // [Ntot][3]
  const int64_t* labels;      // [Ntot]
  const int64_t* faces;       // [npf][Ftot]
  const int64_t* op_row;      // [Ztot] dataset-global node ids (may be null)
  const int64_t* op_col;
  const float* op_val;
  int64_t Ftot;
  // batch descriptor (device): [node_ptr B+1 | face_ptr B+1 | nnz_ptr B+1 | node_src B | face_src B | nnz_src B]
  const int64_t* meta;
  int B, npf;
  int64_t N, F, Z;
  // outputs (device)
  double* o_pos64;
  float* o_pos32;
  float* o_mean_stress;
  float* o_local_stress;
  int64_t* o_labels;
  int64_t* o_batch;
  int64_t* o_faces;   // [npf][F]
  int64_t* o_op_idx;  // [2][Z]: rows, cols (sparse COO indices)
  float* o_op_val;
};

__device__ __forceinline__ int seg_of(const int64_t* __restrict__ ptr, int B, int64_t i) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_resident_gather(ResidentGatherArgs a) {
  const int B = a.B;
  const int64_t* node_ptr = a.meta;
  const int64_t* face_ptr = a.meta + (B + 1);
  const int64_t* nnz_ptr = a.meta + 2 * (B + 1);
  const int64_t* node_src = a.meta + 3 * (B + 1);
  const int64_t* face_src = node_src + B;
  const int64_t* nnz_src = face_src + B;
  const int64_t total = a.N + a.F + a.Z;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < a.N) {
      const int s = seg_of(node_ptr, B, i);
      const int64_t src = node_src[s] + (i - node_ptr[s]);
      const double2 p = reinterpret_cast<const double2*>(a.pos64)[src];
      reinterpret_cast<double2*>(a.o_pos64)[i] = p;
      a.o_pos32[2 * i] = (float)p.x;
      a.o_pos32[2 * i + 1] = (float)p.y;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a.o_mean_stress[3 * i + k] = a.mean_stress[3 * src + k];
        a.o_local_stress[3 * i + k] = a.local_stress[3 * src + k];
      }
      a.o_labels[i] = a.labels[src];
      a.o_batch[i] = s;
    } else if (i < a.N + a.F) {
      const int64_t f = i - a.N;
      const int s = seg_of(face_ptr, B, f);
      const int64_t src = face_src[s] + (f - face_ptr[s]);
      for (int k = 0; k < a.npf; ++k) a.o_faces[(int64_t)k * a.F + f] = a.faces[(int64_t)k * a.Ftot + src];
    } else {
      const int64_t z = i - a.N - a.F;
      const int s = seg_of(nnz_ptr, B, z);
      const int64_t src = nnz_src[s] + (z - nnz_ptr[s]);
      a.o_op_idx[z] = a.op_row[src] - node_src[s] + node_ptr[s];
      a.o_op_idx[a.Z + z] = a.op_col[src];
      a.o_op_val[z] = a.op_val[src];
    }
  }
}

}  // namespace pdg

using namespace pdg;

extern "C" int pdg_resident_gather(const double* pos64, const float* mean_stress, const float* local_stress, const int64_t* labels,
                                   const int64_t* faces, int64_t faces_total, int nodes_per_face, const int64_t* op_row,
                                   const int64_t* op_col, const float* op_val, const int64_t* meta, int n_graphs, int64_t n_nodes,
                                   int64_t n_faces, int64_t n_nnz, double* o_pos64, float* o_pos32, float* o_mean_stress,
                                   float* o_local_stress, int64_t* o_labels, int64_t* o_batch, int64_t* o_faces, int64_t* o_op_idx,
                                   float* o_op_val, void* stream_) {
  if (n_graphs <= 0 || n_nodes <= 0 || n_faces <= 0 || n_nnz < 0) { set_error("pdg_resident_gather: empty batch"); return -1; }
  if (nodes_per_face != 3 && nodes_per_face != 4) { set_error("pdg_resident_gather: nodes_per_face = %d", nodes_per_face); return -1; }
  if (n_nnz > 0 && (op_row == nullptr || o_op_idx == nullptr)) { set_error("pdg_resident_gather: operator arrays missing"); return -1; }
  ResidentGatherArgs a{pos64, mean_stress, local_stress, labels, faces, op_row, op_col, op_val, faces_total, meta, n_graphs, nodes_per_face,
                       n_nodes, n_faces, n_nnz, o_pos64, o_pos32, o_mean_stress, o_local_stress, o_labels, o_batch, o_faces, o_op_idx, o_op_val};
  const int64_t total = n_nodes + n_faces + n_nnz;
  int grid = (int)((total + 255) / 256);
  const int cap = 8 * num_sms();
  if (grid > cap) grid = cap;
  k_resident_gather<<<grid, 256, 0, (cudaStream_t)stream_>>>(a);
  PDG_LAUNCH_CHECK();
  return 0;
}
