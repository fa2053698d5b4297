// agents.cuh -- batched agent front ends on the device: the act() side of the loop for
// thousands of lock-step instances, so that a shared-policy step is kernel launches, not a Python loop.
//
//   k_policy_frap    forward of the reference's FRAP Q-network (agents/mplight.py:43-131, the model of MPLight) over
//                    states.mplight rows + greedy choice among the signal's valid actions (agents/agent.py:46-60)
//   k_policy_random  uniform random green phase per signal (the epsilon = 1 exploration of IDQN / IPPO / MPLight,
//                    agents/pfrl_dqn.py:57-70), Philox keyed by (seed, global instance id, signal, tick)
//
// (MAXPRESSURE / MAXWAVE: k_policy in sim.cu.)  Plain fp32 on the CUDA cores: one row is 13 x 12 pair competitions
// of a 20 x 20 layer -- there is no batched-GEMM shape worth a tensor-core tile here and the argmax has to agree with
// the reference's fp32 Q-values.
#pragma once
#include "sim_kernels.cuh"

namespace rs {

// parameter block of FRAP in one float array (offsets in floats); names and shapes of agents/mplight.py:59-71
enum FrapOff {
  FP_P = 0,            // p.weight                  [2][4]
  FP_DW = 8,           // d.weight                  [4][1]   (demand_shape = 1)
  FP_DB = 12,          // d.bias                    [4]
  FP_LEW = 16,         // lane_embedding.weight     [16][8]
  FP_LEB = 144,        // lane_embedding.bias       [16]
  FP_LCW = 160,        // lane_conv.weight          [20][32]
  FP_LCB = 800,        // lane_conv.bias            [20]
  FP_REW = 820,        // relation_embedding.weight [2][4]
  FP_RCW = 828,        // relation_conv.weight      [20][4]
  FP_RCB = 908,        // relation_conv.bias        [20]
  FP_HLW = 928,        // hidden_layer.weight       [20][20]
  FP_HLB = 1328,       // hidden_layer.bias         [20]
  FP_BMW = 1348,       // before_merge.weight       [1][20]
  FP_BMB = 1368,       // before_merge.bias         [1]
  FP_TOTAL = 1369
};
constexpr int kFrapMaxPairs = 16;
constexpr int kFrapThreads = 128;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// derived, batch-independent tables (computed once per block into shared memory)
struct FrapShared {
  float w[FP_TOTAL];
  float ph_part[2][16];   // lane_embedding over the phase half of its input (+ bias), for "movement in the phase" 0 / 1
  float rel[2][20];       // relation branch for competition-mask values 0 / 1 (mplight.py:116-119)
  float lct[16][40];      // lane_conv.weight transposed: [input u][k]; k < 20: first half of the pair, else second half
                          // (threads with consecutive k read consecutive words instead of one bank)
};
// row strides of the per-block scratch arrays, odd so that threads working on different rows / pairs hit different banks
constexpr int kPdStride = 17;   // [R][12] movement embeddings of 16 floats
constexpr int kFsStride = 41;   // [R][n_pairs] pair embeddings of 40 floats: first (20) | second (20)

// obs: [rows][13] states.mplight (phase index, 12 pressures); pairs [n_pairs][2]; comp [n_pairs][n_pairs - 1] 0/1;
// order [S][n_pairs][2] = (pair index, action) in the reference's evaluation order, -1 terminates (k_policy's table);
// q_out [rows][n_pairs] or null; actions [rows].
__global__ void __launch_bounds__(kFrapThreads) k_policy_frap(const float* __restrict__ obs, int rows, int n_signals,
                                                              const float* __restrict__ params, const int32_t* __restrict__ pairs,
                                                              int n_pairs, const uint8_t* __restrict__ comp,
                                                              const int32_t* __restrict__ order, float* __restrict__ q_out,
                                                              int32_t* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char fsm[];
  FrapShared& F = *reinterpret_cast<FrapShared*>(fsm);
  const int R = kFrapThreads / n_pairs;                 // rows per block
  float* pd = reinterpret_cast<float*>(fsm + sizeof(FrapShared));        // [R][12][kPdStride]
  float* fs = pd + R * 12 * kPdStride;                                   // [R][n_pairs][kFsStride]
  float* qs = fs + R * n_pairs * kFsStride;                              // [R][n_pairs]
  const int tid = threadIdx.x;
  for (int i = tid; i < FP_TOTAL; i += kFrapThreads) F.w[i] = __ldg(params + i);
  for (int x = tid; x < 16 * 40; x += kFrapThreads) {
    const int u = x / 40, k = x % 40;
    F.lct[u][k] = __ldg(params + FP_LCW + (k < 20 ? k : k - 20) * 32 + (k < 20 ? 0 : 16) + u);
  }
  __syncthreads();
  if (tid < 32) {
    const int e = tid >> 4, u = tid & 15;
    float acc = F.w[FP_LEB + u];
    for (int c = 0; c < 4; ++c) acc = __fmaf_rn(F.w[FP_LEW + u * 8 + c], sigmoidf_(F.w[FP_P + e * 4 + c]), acc);
    F.ph_part[e][u] = acc;
  } else if (tid < 72) {
    const int e = (tid - 32) / 20, k = (tid - 32) % 20;
    float acc = F.w[FP_RCB + k];
    for (int c = 0; c < 4; ++c) acc = __fmaf_rn(F.w[FP_RCW + k * 4 + c], fmaxf(F.w[FP_REW + e * 4 + c], 0.0f), acc);
    F.rel[e][k] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  const int row0 = blockIdx.x * R;
  // ---- A: per movement: relu(lane_embedding(cat(sigmoid(p[in phase]), sigmoid(d(x))))) ----
  // the four sigmoid(d(x)) of a movement once (scratch: the pair-embedding array, not yet in use), then the 16 outputs
  float* sg = fs;                                                        // [R][12][4]
  for (int x = tid; x < R * 12 * 4; x += kFrapThreads) {
    const int r = x / 48, m = (x >> 2) % 12, c = x & 3;
    const int row = row0 + r;
    sg[x] = row < rows ? sigmoidf_(__fmaf_rn(F.w[FP_DW + c], __ldg(obs + (size_t)row * 13 + 1 + m), F.w[FP_DB + c])) : 0.0f;
  }
  __syncthreads();
  for (int x = tid; x < R * 12 * 16; x += kFrapThreads) {
    const int r = x / 192, m = (x / 16) % 12, u = x & 15;
    const int row = row0 + r;
    float v = 0.0f;
    if (row < rows) {
      const int act = (int)__ldg(obs + (size_t)row * 13);
      int e = 0;
      if (act >= 0 && act < n_pairs) e = (m == __ldg(pairs + 2 * act) || m == __ldg(pairs + 2 * act + 1)) ? 1 : 0;
      float acc = F.ph_part[e][u];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc = __fmaf_rn(F.w[FP_LEW + u * 8 + 4 + c], sg[(r * 12 + m) * 4 + c], acc);
      v = fmaxf(acc, 0.0f);
    }
    pd[(r * 12 + m) * kPdStride + u] = v;
  }
  __syncthreads();
  // ---- B: pair embeddings through the two halves of the 1x1 lane convolution ----
  for (int x = tid; x < R * n_pairs * 40; x += kFrapThreads) {
    const int r = x / (n_pairs * 40), i = (x / 40) % n_pairs, k = x % 40;
    const float* pa = pd + (r * 12 + __ldg(pairs + 2 * i)) * kPdStride;
    const float* pb = pd + (r * 12 + __ldg(pairs + 2 * i + 1)) * kPdStride;
    float acc = k < 20 ? F.w[FP_LCB + k] : 0.0f;
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = __fmaf_rn(F.lct[u][k], pa[u] + pb[u], acc);
    fs[(r * n_pairs + i) * kFsStride + k] = acc;
  }
  __syncthreads();
  // ---- C: pair competition: one thread per (row, pair i), the n - 1 opponents three at a time so that every weight of
  //      the 20 x 20 layer (one 128-bit shared-memory broadcast load per four) feeds three FMAs ----
  const int r = tid / n_pairs, i = tid % n_pairs;
  if (r < R) {
    const float* first = fs + (r * n_pairs + i) * kFsStride;
    float f[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) f[k] = first[k];
    float q = 0.0f;
    const int n_opp = n_pairs - 1;
    for (int jj = 0; jj < n_opp; jj += 3) {
      float c[3][20];
      bool live[3];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        live[t] = jj + t < n_opp;
        const int jo = live[t] ? jj + t : jj;                       // a dead slot repeats a live one, its result is dropped
        const int j = jo < i ? jo : jo + 1;                         // the jo-th OTHER pair, in the reference's order
        const float* sj = fs + (r * n_pairs + j) * kFsStride + 20;
        const float* rl = F.rel[__ldg(comp + i * n_opp + jo)];
#pragma unroll
        for (int k = 0; k < 20; ++k) c[t][k] = fmaxf(f[k] + sj[k], 0.0f) * rl[k];
      }
      float o[3] = {F.w[FP_BMB], F.w[FP_BMB], F.w[FP_BMB]};
#pragma unroll 2
      for (int k2 = 0; k2 < 20; ++k2) {
        const float4* w4 = reinterpret_cast<const float4*>(F.w + FP_HLW + k2 * 20);
        const float hb = F.w[FP_HLB + k2];
        float h[3] = {hb, hb, hb};
#pragma unroll
        for (int kq = 0; kq < 5; ++kq) {
          const float4 w = w4[kq];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            h[t] = __fmaf_rn(w.x, c[t][4 * kq + 0], h[t]); h[t] = __fmaf_rn(w.y, c[t][4 * kq + 1], h[t]);
            h[t] = __fmaf_rn(w.z, c[t][4 * kq + 2], h[t]); h[t] = __fmaf_rn(w.w, c[t][4 * kq + 3], h[t]);
          }
        }
        const float bm = F.w[FP_BMW + k2];
#pragma unroll
        for (int t = 0; t < 3; ++t) o[t] = __fmaf_rn(bm, fmaxf(h[t], 0.0f), o[t]);
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) if (live[t]) q += o[t];
    }
    qs[r * n_pairs + i] = q;
    const int row = row0 + r;
    if (row < rows && q_out) q_out[(size_t)row * n_pairs + i] = q;
  }
  __syncthreads();
  // ---- greedy action among the signal's valid pairs, first maximum in the reference's scan order ----
  if (tid < R && row0 + tid < rows) {
    const int row = row0 + tid, sg = row % n_signals;
    float best = 0.0f; int bi = -1;
    for (int k = 0; k < n_pairs; ++k) {
      const int p = __ldg(order + (sg * n_pairs + k) * 2);
      if (p < 0) break;
      const float v = qs[tid * n_pairs + p];
      if (bi < 0 || v > best) { best = v; bi = __ldg(order + (sg * n_pairs + k) * 2 + 1); }
    }
    actions[row] = bi < 0 ? 0 : bi;
  }
}

inline size_t frap_smem_bytes(int n_pairs) {
  const int R = kFrapThreads / n_pairs;
  return sizeof(FrapShared) + sizeof(float) * ((size_t)R * 12 * kPdStride + (size_t)R * n_pairs * kFsStride + (size_t)R * n_pairs);
}

// uniform random green phase: Philox keyed by (seed, global instance id, signal, instance tick)
__global__ void k_policy_random(DevSim D, uint64_t seed, int32_t* actions) {
  const DevScenario& sc = D.sc;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= D.n_env * sc.n_signals) return;
  const int env = x / sc.n_signals, sg = x % sc.n_signals;
  const uint64_t id = (uint64_t)(D.first_env_id + env);
  uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)sg, (uint32_t)D.hdr[(size_t)env * kHdrInts + H_TICK]};
  philox(c, (uint32_t)seed ^ 0x5EED5EEDu, (uint32_t)(seed >> 32));
  actions[x] = (int32_t)(c[0] % (uint32_t)__ldg(sc.sig_n_green + sg));
}

}  // namespace rs
