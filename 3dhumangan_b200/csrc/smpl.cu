// SMPL linear blend skinning: the step in front of the hot path that produces the generator's pose conditions
// (SURVEY.md 8f-4).  Reference: `lbs` (lib/components/smpl.py:11-107), which composes smplx.lbs' blend_shapes,
// vertices2joints, batch_rodrigues and batch_rigid_transform (smplx is not vendored by the reference; their published
// algorithm is restated in oracle/smpl_port.py), and the skinning of `SHHQDataset._preprocess_smpl_fix_body`
// (lib/data/datasets.py:146-155).  Three small launches per batch:
//   hg_smpl_shape  v_shaped = v_template + shapedirs . betas;  per-block partials of J = J_regressor . v_shaped
//   hg_smpl_pose   per sample: J (fixed-order sum of the partials), Rodrigues, pose feature (R - I), kinematic chain,
//                  A_j = T_j - [0 | T_j J_j]  (the "fk_matrices" the generator consumes), posed joints
//   hg_smpl_skin   per vertex: pose blend shapes (207 x 3 MACs), T_v = sum_j w[v,j] A_j, vertex = T_v [v_posed; 1]
// fp32 throughout, deterministic.  V = 6890, J = 24 for SMPL, but nothing here depends on those numbers (J <= 32).
#include "common.cuh"

namespace hg {

constexpr int kSmplMaxJ = 32;

__global__ void __launch_bounds__(256) smpl_shape_kernel(const float* __restrict__ v_template, const float* __restrict__ shapedirs,
                                                         const float* __restrict__ betas, const float* __restrict__ jreg,
                                                         float* __restrict__ v_shaped, float* __restrict__ jpart, int V, int NB,
                                                         int J) {
  __shared__ float sb[64];
  __shared__ float red[8][3];
  const int b = blockIdx.y, v = blockIdx.x * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < NB; i += 256) sb[i] = betas[static_cast<long>(b) * NB + i];
  __syncthreads();
  float p[3] = {0.f, 0.f, 0.f};
  if (v < V) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = v_template[v * 3 + c];
      const float* sd = shapedirs + (static_cast<long>(v) * 3 + c) * NB;
      for (int l = 0; l < NB; ++l) acc = fmaf(sb[l], sd[l], acc);
      p[c] = acc;
      v_shaped[(static_cast<long>(b) * V + v) * 3 + c] = acc;
    }
  }
  // partial joint regression of this block's 256 vertices: jpart[b][block][j][c]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < J; ++j) {
    const float w = v < V ? jreg[static_cast<long>(j) * V + v] : 0.f;
    float s[3] = {w * p[0], w * p[1], w * p[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    if (lane == 0) { red[warp][0] = s[0]; red[warp][1] = s[1]; red[warp][2] = s[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
      jpart[((static_cast<long>(b) * gridDim.x + blockIdx.x) * J + j) * 3 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void mat4_mul(const float* a, const float* b, float* o) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s = fmaf(a[i * 4 + k], b[k * 4 + j], s);
      o[i * 4 + j] = s;
    }
}

// one block of 32 threads per sample
__global__ void __launch_bounds__(32) smpl_pose_kernel(const float* __restrict__ jpart, int nblk, const float* __restrict__ pose,
                                                       int pose_is_rotmat, const int* __restrict__ parents,
                                                       float* __restrict__ J_out, float* __restrict__ rot_out,
                                                       float* __restrict__ feat_out, float* __restrict__ A_out,
                                                       float* __restrict__ Jt_out, int J) {
  __shared__ float sJ[kSmplMaxJ][3], sR[kSmplMaxJ][9], sT[kSmplMaxJ][16];
  const int b = blockIdx.x, j = threadIdx.x;
  if (j < J) {
    for (int c = 0; c < 3; ++c) {
      float t = 0.f;
      for (int k = 0; k < nblk; ++k) t += jpart[((static_cast<long>(b) * nblk + k) * J + j) * 3 + c];      // fixed order
      sJ[j][c] = t;
      J_out[(static_cast<long>(b) * J + j) * 3 + c] = t;
    }
    float R[9];
    if (pose_is_rotmat) {
      for (int i = 0; i < 9; ++i) R[i] = pose[(static_cast<long>(b) * J + j) * 9 + i];
    } else {
      // smplx.lbs.batch_rodrigues: angle = |r + 1e-8|, axis = r / angle, R = I + sin K + (1 - cos) K^2
      const float* r = pose + (static_cast<long>(b) * J + j) * 3;
      const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
      const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
      const float rx = r[0] / angle, ry = r[1] / angle, rz = r[2] / angle;
      const float s = sinf(angle), c = cosf(angle), oc = 1.f - c;
      // K = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]];  K^2 = r r^T - I (|r| = 1 up to rounding: computed as the product)
      const float K[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
      float K2[9];
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) {
          float t = 0.f;
          for (int m = 0; m < 3; ++m) t = fmaf(K[i * 3 + m], K[m * 3 + k], t);
          K2[i * 3 + k] = t;
        }
      for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.f : 0.f) + s * K[i] + oc * K2[i];
    }
    for (int i = 0; i < 9; ++i) {
      sR[j][i] = R[i];
      rot_out[(static_cast<long>(b) * J + j) * 9 + i] = R[i];
      if (j >= 1) feat_out[static_cast<long>(b) * (J - 1) * 9 + (j - 1) * 9 + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
    }
  }
  __syncwarp();
  if (j == 0) {   // kinematic chain (smplx.lbs.batch_rigid_transform): T_i = T_parent(i) . [R_i | J_i - J_parent(i)]
    for (int i = 0; i < J; ++i) {
      const int pa = parents[i];
      float rel[3];
      for (int c = 0; c < 3; ++c) rel[c] = sJ[i][c] - (i > 0 ? sJ[pa][c] : 0.f);
      float M[16] = {sR[i][0], sR[i][1], sR[i][2], rel[0], sR[i][3], sR[i][4], sR[i][5], rel[1],
                     sR[i][6], sR[i][7], sR[i][8], rel[2], 0.f, 0.f, 0.f, 1.f};
      if (i == 0) {
        for (int k = 0; k < 16; ++k) sT[0][k] = M[k];
      } else {
        mat4_mul(sT[pa], M, sT[i]);
      }
    }
  }
  __syncwarp();
  if (j < J) {
    float* A = A_out + (static_cast<long>(b) * J + j) * 16;
    // A = T - [0 | T . [J;0]]  (rel_transforms: removes the rest-pose joint location)
    float tj[4];
    for (int i = 0; i < 4; ++i) tj[i] = sT[j][i * 4 + 0] * sJ[j][0] + sT[j][i * 4 + 1] * sJ[j][1] + sT[j][i * 4 + 2] * sJ[j][2];
    for (int i = 0; i < 4; ++i)
      for (int k = 0; k < 4; ++k) A[i * 4 + k] = sT[j][i * 4 + k] - (k == 3 ? tj[i] : 0.f);
    for (int c = 0; c < 3; ++c) Jt_out[(static_cast<long>(b) * J + j) * 3 + c] = sT[j][c * 4 + 3];
  }
}

// verts[b,v] = (sum_j w[v,j] A[b,j]) . [v_in[b,v] + posedirs^T feat[b]; 1]      (feat null: no pose blend shapes)
__global__ void __launch_bounds__(128) smpl_skin_kernel(const float* __restrict__ v_in, long v_bstride, const float* __restrict__ feat,
                                                        const float* __restrict__ posedirs, int P, const float* __restrict__ lbsw,
                                                        long w_bstride, const float* __restrict__ A, float* __restrict__ verts,
                                                        int V, int J) {
  __shared__ float sA[kSmplMaxJ * 16];
  extern __shared__ float sfeat[];
  const int b = blockIdx.y, v = blockIdx.x * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < J * 16; i += 128) sA[i] = A[static_cast<long>(b) * J * 16 + i];
  if (feat)
    for (int i = threadIdx.x; i < P; i += 128) sfeat[i] = feat[static_cast<long>(b) * P + i];
  __syncthreads();
  if (v >= V) return;
  float p[3];
  for (int c = 0; c < 3; ++c) p[c] = v_in[b * v_bstride + static_cast<long>(v) * 3 + c];
  if (feat) {
    float o[3] = {0.f, 0.f, 0.f};
    for (int q = 0; q < P; ++q) {
      const float* pd = posedirs + static_cast<long>(q) * V * 3 + v * 3;
      const float f = sfeat[q];
      o[0] = fmaf(f, pd[0], o[0]);
      o[1] = fmaf(f, pd[1], o[1]);
      o[2] = fmaf(f, pd[2], o[2]);
    }
    p[0] += o[0]; p[1] += o[1]; p[2] += o[2];
  }
  float T[12];
  for (int i = 0; i < 12; ++i) T[i] = 0.f;
  const float* w = lbsw + b * w_bstride + static_cast<long>(v) * J;
  for (int j = 0; j < J; ++j) {
    const float wj = w[j];
    for (int i = 0; i < 12; ++i) T[i] = fmaf(wj, sA[j * 16 + i], T[i]);
  }
  for (int c = 0; c < 3; ++c)
    verts[(static_cast<long>(b) * V + v) * 3 + c] = T[c * 4 + 0] * p[0] + T[c * 4 + 1] * p[1] + T[c * 4 + 2] * p[2] + T[c * 4 + 3];
}

}  // namespace hg

extern "C" {

int hg_smpl_shape_blocks(int V) { return (V + 255) / 256; }

// v_shaped [B,V,3]; jpart [B, hg_smpl_shape_blocks(V), J, 3] workspace
int hg_smpl_shape(const float* v_template, const float* shapedirs, const float* betas, const float* j_regressor, float* v_shaped,
                  float* jpart, int B, int V, int NB, int J, void* stream) {
  HG_REQUIRE(v_template && shapedirs && betas && j_regressor && v_shaped && jpart, "hg_smpl_shape: null pointer");
  HG_REQUIRE(B > 0 && V > 0 && NB >= 0 && NB <= 64 && J >= 1 && J <= hg::kSmplMaxJ, "hg_smpl_shape: bad sizes (betas <= 64, joints <= 32)");
  dim3 grid((V + 255) / 256, B);
  hg::smpl_shape_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(v_template, shapedirs, betas, j_regressor, v_shaped, jpart, V,
                                                                            NB, J);
  return hg::check_launch("hg_smpl_shape");
}

// pose: [B,J,3] axis-angle (pose_is_rotmat = 0) or [B,J,9] rotation matrices.  Outputs: J [B,J,3], rot [B,J,9], feat [B,(J-1)*9],
// A [B,J,16] (rest-pose-relative rigid transforms = the reference's fk_matrices), posed joints [B,J,3].
int hg_smpl_pose(const float* jpart, int nblk, const float* pose, int pose_is_rotmat, const int* parents, float* joints, float* rot,
                 float* feat, float* A, float* joints_posed, int B, int J, void* stream) {
  HG_REQUIRE(jpart && pose && parents && joints && rot && feat && A && joints_posed, "hg_smpl_pose: null pointer");
  HG_REQUIRE(B > 0 && J >= 1 && J <= hg::kSmplMaxJ && nblk > 0, "hg_smpl_pose: bad sizes");
  hg::smpl_pose_kernel<<<B, 32, 0, static_cast<cudaStream_t>(stream)>>>(jpart, nblk, pose, pose_is_rotmat, parents, joints, rot, feat, A,
                                                                       joints_posed, J);
  return hg::check_launch("hg_smpl_pose");
}

// verts [B,V,3] = skin(v_in (+ posedirs^T feat), lbs weights, A).  v_in / lbs weights may be shared by the batch (stride 0).
int hg_smpl_skin(const float* v_in, long v_bstride, const float* feat, const float* posedirs, int P, const float* lbs_weights,
                 long w_bstride, const float* A, float* verts, int B, int V, int J, void* stream) {
  HG_REQUIRE(v_in && lbs_weights && A && verts, "hg_smpl_skin: null pointer");
  HG_REQUIRE((feat == nullptr) == (posedirs == nullptr), "hg_smpl_skin: feat and posedirs go together");
  HG_REQUIRE(B > 0 && V > 0 && J >= 1 && J <= hg::kSmplMaxJ && P >= 0 && P <= 4096, "hg_smpl_skin: bad sizes");
  dim3 grid((V + 127) / 128, B);
  hg::smpl_skin_kernel<<<grid, 128, static_cast<size_t>(P) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      v_in, v_bstride, feat, posedirs, P, lbs_weights, w_bstride, A, verts, V, J);
  return hg::check_launch("hg_smpl_skin");
}

}  // extern "C"
