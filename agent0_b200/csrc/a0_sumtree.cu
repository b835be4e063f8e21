// GPU sum-tree for prioritized replay: K2a (batched stratified draws + IS weights) and K2b (leaf
// writes with bottom-up recomputation).
//
// Replaces the reference's flat CPU priority vector: the intended draw torch.multinomial(priority)
// (agent0/deepq/replay.py:39-43), priority[ids] = (loss+eps)^alpha and max_p (replay.py:55-59),
// priority[-k:] = max_p**alpha on extend (replay.py:52) and the IS-weight block of Trainer.step
// (agent0/deepq/trainer.py:91-94), which sums 1 M CPU floats per update.
//
// Arithmetic is defined by oracle/sumtree.py and must match it bit for bit:
//   node = fl32(left + right), recomputed from children;  t = fl32(fl32(fl32(b + u)/B) * root);
//   go left iff (t < left) || !(right > 0);  going right: t = fl32(t - left).
// The tree (2P floats, 8 MB at 1 M leaves, 16.8 MB at 2 M) is L2-resident: these kernels are
// latency-bound, not HBM-bound.
#include "a0_common.cuh"

// ------------------------------------------------------------------------------------------------
// K2a.  One warp per draw.  Instead of one dependent L2 round trip per level, the warp fetches the
// whole 5-level sub-heap below the current node in two loads (lane i: sub-heap node 32+i; lane h-2:
// sub-heap node h for h = 2..31) and walks it through shuffles, with exactly the scalar arithmetic.
// ------------------------------------------------------------------------------------------------
constexpr int K2A_WARPS = 8;

__global__ void __launch_bounds__(K2A_WARPS * 32)
a0_k2a_sample(const float* __restrict__ tree, int64_t P, int32_t D, const float* __restrict__ u, int32_t total,
              int32_t batch, float top, float beta, float sum_offset, int32_t uniform,
              int64_t* __restrict__ idx_out, float* __restrict__ prio_out, float* __restrict__ weight_out,
              unsigned int* counter) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = blockIdx.x * K2A_WARPS + warp;
  const float root = __ldcg(tree + 1);
  if (g < total) {
    const int b = g % batch;
    float t = __fmul_rn(__fdiv_rn(__fadd_rn((float)b, u[g]), (float)batch), root);
    int64_t v = 1;          // current node, warp-uniform
    float leaf = root;
    int left = D;
    while (left > 0) {
      const int c = left < 5 ? left : 5;
      // sub-heap node h lives at tree[(v << l) + (h - (1 << l))], l = floor(log2 h)
      const int hB = lane + 2;
      const int lB = 31 - __clz(hB);
      float regB = 0.0f, regA = 0.0f;
      if (lB <= c && lB <= 4 && hB < 32) regB = __ldcg(tree + ((v << lB) + (hB - (1 << lB))));
      if (c == 5) regA = __ldcg(tree + ((v << 5) + lane));
      int h = 1;
      for (int s = 0; s < c; ++s) {
        const int hl = 2 * h;
        float L, R;
        if (hl < 32) {
          L = __shfl_sync(0xffffffffu, regB, hl - 2);
          R = __shfl_sync(0xffffffffu, regB, hl - 1);
        } else {
          L = __shfl_sync(0xffffffffu, regA, hl - 32);
          R = __shfl_sync(0xffffffffu, regA, hl - 31);
        }
        if (t < L || !(R > 0.0f)) { h = hl; leaf = L; }
        else { t = __fsub_rn(t, L); h = hl + 1; leaf = R; }
      }
      v = (v << c) + (h - (1 << c));
      left -= c;
    }
    if (lane == 0) {
      idx_out[g] = v - P;
      prio_out[g] = leaf;
    }
  }
  if (weight_out == nullptr || g >= total) return;
  // ---- epilogue: the last warp of each batch turns its priorities into normalised IS weights
  //      (trainer.py:91-94), so all batches of a multi-batch draw finish in parallel -------------
  const int k = g / batch;
  unsigned ticket = 0;
  if (lane == 0) {
    __threadfence();
    ticket = atomicAdd(counter + k, 1u);
  }
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != (unsigned)batch - 1u) return;
  __threadfence();
  const float* p = prio_out + (size_t)k * batch;
  float* w = weight_out + (size_t)k * batch;
  if (uniform) {
    for (int j = lane; j < batch; j += 32) w[j] = 1.0f;
  } else {
    const float denom = root + sum_offset;
    float mx = 0.0f;
    for (int j = lane; j < batch; j += 32) {
      const float wj = powf(__fmul_rn(top, __fdiv_rn(__ldcg(p + j), denom)), -beta);
      w[j] = wj;
      mx = fmaxf(mx, wj);
    }
    mx = a0_warp_max(mx);
    const float inv = mx + 1e-8f;
    __syncwarp();
    for (int j = lane; j < batch; j += 32) w[j] = __fdiv_rn(w[j], inv);
  }
  if (lane == 0) counter[k] = 0u;
}

extern "C" int a0_pt_sample(a0_replay_t* h, const float* u, int32_t total, int32_t batch, float top, float beta,
                            float sum_offset, int32_t uniform, int64_t* idx_out, float* prio_out,
                            float* weight_out, a0_stream_t stream_) {
  A0_REQUIRE(h != nullptr, "a0_pt_sample: handle is NULL");
  A0_REQUIRE(total >= 0 && batch > 0 && total % batch == 0, "a0_pt_sample: total %d must be a multiple of batch %d", total, batch);
  if (total == 0) return A0_OK;
  A0_REQUIRE(u && idx_out && prio_out, "a0_pt_sample: NULL argument");
  A0_REQUIRE(total / batch <= A0_MAX_BATCHES, "a0_pt_sample: at most %d batches per call", A0_MAX_BATCHES);
  A0DeviceGuard guard(h->device);
  const int blocks = (total + K2A_WARPS - 1) / K2A_WARPS;
  a0_k2a_sample<<<blocks, K2A_WARPS * 32, 0, (cudaStream_t)stream_>>>(
      h->tree, h->P, h->D, u, total, batch, top, beta, sum_offset, uniform, idx_out, prio_out, weight_out, h->counter);
  A0_LAUNCH_CHECK();
  return A0_OK;
}

// ------------------------------------------------------------------------------------------------
// K2b.  Two launches.
//  (1) a0_k2b_write, one CTA: leaves are written with a deterministic last-writer-wins rule and the
//      4096-leaf chunks they fall in are flagged dirty.
//        mode 0: value = (loss+eps)^alpha, skipped when the leaf is 0; max_p = max(max_p, max loss)
//        mode 1: pos >= 0 -> max_p^alpha;  pos < 0 -> leaf ~pos = 0
//        mode 2: value = vals[k]
//  (2) a0_k2b_rebuild, one CTA per chunk: a dirty chunk's sub-tree is recomputed bottom-up in shared
//      memory, every node as fl32(left + right); the last CTA to finish recomputes the levels above
//      the chunk roots.  The work does not depend on how many leaves changed (a 1 M-leaf tree is
//      4 MB read + 4 MB written, all L2-resident) and the result is bit-reproducible: no atomics
//      on values, no ordering dependence.
// ------------------------------------------------------------------------------------------------
constexpr int K2B_THREADS = 1024;

__global__ void __launch_bounds__(K2B_THREADS)
a0_k2b_write(float* __restrict__ tree, int64_t P, int32_t chunk_log, int64_t N, const int64_t* __restrict__ idx64,
             const int32_t* __restrict__ idx32, const float* __restrict__ vals, int32_t count, int32_t mode,
             float alpha, float eps, float* __restrict__ max_p, int32_t* __restrict__ winner,
             int32_t* __restrict__ dirty) {
  __shared__ float red[K2B_THREADS / 32];
  const int tid = threadIdx.x;
  const float maxp_in = *max_p;
  auto position = [&](int k) -> int64_t {
    if (mode == 1) { const int32_t p = idx32[k]; return p >= 0 ? p : ~p; }
    return idx64[k];
  };
  float local_max = 0.0f;
  for (int k = tid; k < count; k += K2B_THREADS) {          // claim
    const int64_t pos = position(k);
    if (pos < 0 || pos >= N) continue;
    atomicMax(winner + pos, k);
    if (mode == 0) local_max = fmaxf(local_max, vals[k]);
  }
  __syncthreads();
  for (int k = tid; k < count; k += K2B_THREADS) {          // write
    const int64_t pos = position(k);
    if (pos < 0 || pos >= N) continue;
    if (__ldcg(winner + pos) != k) continue;
    float v;
    if (mode == 0) {
      if (!(__ldcg(tree + P + pos) > 0.0f)) continue;      // evicted since it was sampled
      v = a0_priority(vals[k], eps, alpha);
    } else if (mode == 1) {
      v = idx32[k] >= 0 ? (alpha == 0.5f ? sqrtf(maxp_in) : powf(maxp_in, alpha)) : 0.0f;
    } else {
      v = vals[k];
    }
    tree[P + pos] = v;
    dirty[pos >> chunk_log] = 1;
  }
  __syncthreads();
  for (int k = tid; k < count; k += K2B_THREADS) {          // release
    const int64_t pos = position(k);
    if (pos >= 0 && pos < N) winner[pos] = -1;
  }
  if (mode == 0) {
    local_max = a0_warp_max(local_max);
    if ((tid & 31) == 0) red[tid >> 5] = local_max;
    __syncthreads();
    if (tid == 0) {
      float mx = maxp_in;
      for (int i = 0; i < K2B_THREADS / 32; ++i) mx = fmaxf(mx, red[i]);
      *max_p = mx;
    }
  }
}

constexpr int K2R_THREADS = 512;
constexpr int K2R_MAXLOG = 12;      // chunk of up to 4096 leaves; up to 4096 chunk roots above

// Reduce `n = 1 << levels` values held in buf[0][0..n) up to one value, writing every level to the
// tree: the node at `levels - l` levels above the inputs, i-th of its row, goes to tree[base(l) + i].
__device__ __forceinline__ void a0_reduce_levels(float (*buf)[1 << K2R_MAXLOG], int levels, float* tree,
                                                 int64_t row0, int64_t first) {
  // inputs sit at tree row `row0 + levels` (depth), starting at column `first << levels`
  int cur = 0;
  for (int l = levels - 1; l >= 0; --l) {
    const int cnt = 1 << l;
    const int64_t base = ((int64_t)1 << (row0 + l)) + (first << l);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float v = __fadd_rn(buf[cur][2 * i], buf[cur][2 * i + 1]);
      buf[cur ^ 1][i] = v;
      tree[base + i] = v;
    }
    __syncthreads();
    cur ^= 1;
  }
}

__global__ void __launch_bounds__(K2R_THREADS)
a0_k2b_rebuild(float* __restrict__ tree, int32_t D, int32_t chunk_log, int32_t* __restrict__ dirty,
               unsigned int* __restrict__ ticket) {
  __shared__ float buf[2][1 << K2R_MAXLOG];
  __shared__ bool is_last;
  const int c = blockIdx.x;
  const int top_levels = D - chunk_log;                  // depth of the chunk roots
  if (__ldcg(dirty + c) != 0) {
    const int n = 1 << chunk_log;
    const float* leaves = tree + ((int64_t)1 << D) + ((int64_t)c << chunk_log);
    for (int i = threadIdx.x; i < n; i += blockDim.x) buf[0][i] = __ldcg(leaves + i);
    __syncthreads();
    a0_reduce_levels(buf, chunk_log, tree, top_levels, c);
    if (threadIdx.x == 0) dirty[c] = 0;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (top_levels > 0) {
    const int n = 1 << top_levels;
    const float* roots = tree + ((int64_t)1 << top_levels);
    for (int i = threadIdx.x; i < n; i += blockDim.x) buf[0][i] = __ldcg(roots + i);
    __syncthreads();
    a0_reduce_levels(buf, top_levels, tree, 0, 0);
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

static int a0_launch_update(a0_replay_t* h, const int64_t* idx64, const int32_t* idx32, const float* vals,
                            int32_t count, int32_t mode, float alpha, float eps, a0_stream_t stream_) {
  if (count == 0) return A0_OK;
  A0DeviceGuard guard(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int chunk_log = h->D < K2R_MAXLOG ? h->D : K2R_MAXLOG;
  a0_k2b_write<<<1, K2B_THREADS, 0, stream>>>(h->tree, h->P, chunk_log, h->N, idx64, idx32, vals, count, mode, alpha,
                                             eps, h->max_p, h->winner, h->dirty);
  A0_LAUNCH_CHECK();
  a0_k2b_rebuild<<<(unsigned)(h->P >> chunk_log), K2R_THREADS, 0, stream>>>(h->tree, h->D, chunk_log, h->dirty,
                                                                           h->counter + A0_MAX_BATCHES);
  A0_LAUNCH_CHECK();
  return A0_OK;
}

extern "C" int a0_pt_update(a0_replay_t* h, const int64_t* idx, const float* loss, int32_t count, float alpha,
                            float eps, a0_stream_t stream) {
  A0_REQUIRE(h != nullptr && count >= 0, "a0_pt_update: bad handle or count");
  A0_REQUIRE(count == 0 || (idx && loss), "a0_pt_update: NULL argument");
  return a0_launch_update(h, idx, nullptr, loss, count, 0, alpha, eps, stream);
}
extern "C" int a0_pt_mark(a0_replay_t* h, const int32_t* pos, int32_t count, float alpha, a0_stream_t stream) {
  A0_REQUIRE(h != nullptr && count >= 0, "a0_pt_mark: bad handle or count");
  A0_REQUIRE(count == 0 || pos, "a0_pt_mark: NULL argument");
  return a0_launch_update(h, nullptr, pos, nullptr, count, 1, alpha, 0.0f, stream);
}
extern "C" int a0_pt_set(a0_replay_t* h, const int64_t* idx, const float* value, int32_t count, a0_stream_t stream) {
  A0_REQUIRE(h != nullptr && count >= 0, "a0_pt_set: bad handle or count");
  A0_REQUIRE(count == 0 || (idx && value), "a0_pt_set: NULL argument");
  return a0_launch_update(h, idx, nullptr, value, count, 2, 1.0f, 0.0f, stream);
}
