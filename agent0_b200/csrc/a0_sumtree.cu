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
  A0_PDL_PROLOGUE();
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
  A0_LAUNCH(a0_k2a_sample, (unsigned)blocks, K2A_WARPS * 32, 0, (cudaStream_t)stream_, 1, h->tree, h->P, h->D, u, total, batch,
            top, beta, sum_offset, uniform, idx_out, prio_out, weight_out, h->counter);
  return A0_OK;
}

// ------------------------------------------------------------------------------------------------
// K2b.  Leaves are written with a deterministic last-writer-wins rule:
//        mode 0: value = (loss+eps)^alpha, skipped when the leaf is 0; max_p = max(max_p, max loss)
//        mode 1: pos >= 0 -> max_p^alpha;  pos < 0 -> leaf ~pos = 0
//        mode 2: value = vals[k]
// and every touched ancestor is recomputed as fl32(left + right) -- never a delta, so the tree is
// bit-reproducible whatever the order of the updates.  Two implementations, same result:
//  * count <= K2P_MAX = 16384 (the training loop: B..20*B sampled indices, a few hundred marks per
//    append): a0_k2b_paths, ONE launch of one thread-block cluster (1..8 CTAs x 1024 threads, at
//    most two indices per thread, held in registers); the phases are separated by the hardware
//    cluster barrier.  After the leaf writes every index walks to the root
//    three levels per step: a thread loads the 8 siblings under its great-grandparent (two 16-byte
//    L2 loads), forms the 4 + 2 + 1 nodes above them and stores them; threads that share
//    ancestors compute identical values from identical inputs, so the duplicate stores are
//    benign.  7 dependent L2 round trips for a 2 M-leaf tree instead of 21.
//  * larger counts (bulk fills): a0_k2b_write (one CTA) marks the 4096-leaf chunks it touched and
//    a0_k2b_rebuild (one CTA per chunk) recomputes dirty chunks bottom-up in shared memory; the
//    last CTA to finish recomputes the levels above the chunk roots.  Work independent of count.
// ------------------------------------------------------------------------------------------------
constexpr int K2B_THREADS = 1024;

__global__ void __launch_bounds__(K2B_THREADS)
a0_k2b_write(float* __restrict__ tree, int64_t P, int32_t chunk_log, int64_t N, const int64_t* __restrict__ idx64,
             const int32_t* __restrict__ idx32, const float* __restrict__ vals, int32_t count, int32_t mode,
             float alpha, float eps, float* __restrict__ max_p, int32_t* __restrict__ winner,
             int32_t* __restrict__ dirty) {
  __shared__ float red[K2B_THREADS / 32];
  A0_PDL_PROLOGUE();
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
  A0_PDL_PROLOGUE();
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

constexpr int K2P_THREADS = 1024;
constexpr int K2P_CLUSTER = 8;                      // portable cluster size
constexpr int K2P_PER_THREAD = 2;
constexpr int K2P_MAX = K2P_THREADS * K2P_CLUSTER * K2P_PER_THREAD;   // 16384 indices per launch

// All CTAs of the launch form ONE thread-block cluster, so the hardware cluster barrier is a
// grid-wide barrier; release/acquire at cluster scope orders the L2-level (ld.cg/st.cg) accesses.
__device__ __forceinline__ void a0_cluster_sync(bool single_cta) {
  if (single_cta) { __syncthreads(); return; }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(K2P_THREADS)
a0_k2b_paths(float* __restrict__ tree, int64_t P, int32_t D, int64_t N, const int64_t* __restrict__ idx64,
             const int32_t* __restrict__ idx32, const float* __restrict__ vals, int32_t count, int32_t mode,
             float alpha, float eps, float* __restrict__ max_p, int32_t* __restrict__ winner) {
  A0_PDL_PROLOGUE();
  const bool single = gridDim.x == 1;
  const int gtid = blockIdx.x * K2P_THREADS + threadIdx.x;
  const int gstride = gridDim.x * K2P_THREADS;
  const float maxp_in = __ldcg(max_p);
  // this thread's (at most K2P_PER_THREAD) indices, kept in registers for all phases
  int64_t pos[K2P_PER_THREAD];
  int kk[K2P_PER_THREAD];
  float val[K2P_PER_THREAD];
#pragma unroll
  for (int s = 0; s < K2P_PER_THREAD; ++s) {
    const int k = gtid + s * gstride;
    kk[s] = k;
    pos[s] = -1;
    val[s] = 0.0f;
    if (k < count) {
      int64_t p;
      bool set = true;
      if (mode == 1) { const int32_t q = idx32[k]; set = q >= 0; p = set ? q : ~q; }
      else p = idx64[k];
      if (p >= 0 && p < N) {
        pos[s] = p;
        if (mode == 0) val[s] = vals[k];
        else if (mode == 1) val[s] = set ? 1.0f : 0.0f;
        else val[s] = vals[k];
      }
    }
  }
#pragma unroll
  for (int s = 0; s < K2P_PER_THREAD; ++s)                   // claim
    if (pos[s] >= 0) atomicMax(winner + pos[s], kk[s]);
  if (mode == 0) {
    float mx = 0.0f;
#pragma unroll
    for (int s = 0; s < K2P_PER_THREAD; ++s) mx = fmaxf(mx, pos[s] >= 0 ? val[s] : 0.0f);
    mx = a0_warp_max(mx);
    if ((threadIdx.x & 31) == 0) a0_atomic_max_pos(max_p, mx);   // max_p = max(max_p, max loss), replay.py:59
  }
  a0_cluster_sync(single);
#pragma unroll
  for (int s = 0; s < K2P_PER_THREAD; ++s) {                 // write (the highest k wins)
    if (pos[s] < 0 || __ldcg(winner + pos[s]) != kk[s]) continue;
    float v;
    if (mode == 0) {
      if (!(__ldcg(tree + P + pos[s]) > 0.0f)) continue;    // evicted since it was sampled
      v = a0_priority(val[s], eps, alpha);
    } else if (mode == 1) {
      v = val[s] != 0.0f ? (alpha == 0.5f ? sqrtf(maxp_in) : powf(maxp_in, alpha)) : 0.0f;
    } else {
      v = val[s];
    }
    __stcg(tree + P + pos[s], v);
  }
  a0_cluster_sync(single);
#pragma unroll
  for (int s = 0; s < K2P_PER_THREAD; ++s)                   // release
    if (pos[s] >= 0) winner[pos[s]] = -1;
  // ---- propagate: `depth` is the level of the freshly written nodes, root = level 0 ----------------
  int depth = D;
  int shift = 0;                 // node of an index at `depth` = (P + pos) >> shift
  const int first = D % 3;       // 1 or 2 levels first, then whole groups of three
  if (first == 1) {
#pragma unroll
    for (int s = 0; s < K2P_PER_THREAD; ++s) {
      if (pos[s] < 0) continue;
      const int64_t x = (P + pos[s]) & ~(int64_t)1;
      const float2 c2 = __ldcg(reinterpret_cast<const float2*>(tree + x));
      __stcg(tree + (x >> 1), __fadd_rn(c2.x, c2.y));
    }
  } else if (first == 2) {
#pragma unroll
    for (int s = 0; s < K2P_PER_THREAD; ++s) {
      if (pos[s] < 0) continue;
      const int64_t x = (P + pos[s]) & ~(int64_t)3;
      const float4 c4 = __ldcg(reinterpret_cast<const float4*>(tree + x));
      const float p0 = __fadd_rn(c4.x, c4.y), p1 = __fadd_rn(c4.z, c4.w);
      __stcg(reinterpret_cast<float2*>(tree + (x >> 1)), make_float2(p0, p1));
      __stcg(tree + (x >> 2), __fadd_rn(p0, p1));
    }
  }
  depth -= first;
  shift += first;
  if (first && depth > 0) a0_cluster_sync(single);
  while (depth > 0) {            // depth is a multiple of 3 here
    float4 lo4[K2P_PER_THREAD], hi4[K2P_PER_THREAD];
    int64_t x[K2P_PER_THREAD];
#pragma unroll
    for (int s = 0; s < K2P_PER_THREAD; ++s) {
      x[s] = ((P + (pos[s] < 0 ? 0 : pos[s])) >> shift) & ~(int64_t)7;
      if (pos[s] >= 0) {
        lo4[s] = __ldcg(reinterpret_cast<const float4*>(tree + x[s]));
        hi4[s] = __ldcg(reinterpret_cast<const float4*>(tree + x[s] + 4));
      }
    }
#pragma unroll
    for (int s = 0; s < K2P_PER_THREAD; ++s) {
      if (pos[s] < 0) continue;
      const float p0 = __fadd_rn(lo4[s].x, lo4[s].y), p1 = __fadd_rn(lo4[s].z, lo4[s].w);
      const float p2 = __fadd_rn(hi4[s].x, hi4[s].y), p3 = __fadd_rn(hi4[s].z, hi4[s].w);
      const float g0 = __fadd_rn(p0, p1), g1 = __fadd_rn(p2, p3);
      __stcg(reinterpret_cast<float4*>(tree + (x[s] >> 1)), make_float4(p0, p1, p2, p3));
      __stcg(reinterpret_cast<float2*>(tree + (x[s] >> 2)), make_float2(g0, g1));
      __stcg(tree + (x[s] >> 3), __fadd_rn(g0, g1));
    }
    depth -= 3;
    shift += 3;
    if (depth > 0) a0_cluster_sync(single);
  }
}

static int a0_launch_paths(a0_replay_t* h, const int64_t* idx64, const int32_t* idx32, const float* vals,
                           int32_t count, int32_t mode, float alpha, float eps, cudaStream_t stream) {
  int ctas = (count + K2P_THREADS - 1) / K2P_THREADS;
  ctas = ctas < 1 ? 1 : (ctas > K2P_CLUSTER ? K2P_CLUSTER : ctas);
  if (ctas > 1 && (ctas & (ctas - 1))) {            // cluster sizes: powers of two
    int p2 = 1;
    while (p2 < ctas) p2 <<= 1;
    ctas = p2;
  }
  A0_LAUNCH(a0_k2b_paths, (unsigned)ctas, K2P_THREADS, 0, stream, (unsigned)ctas, h->tree, h->P, h->D, h->N, idx64, idx32, vals,
            count, mode, alpha, eps, h->max_p, h->winner);
  return A0_OK;
}

static int a0_launch_update(a0_replay_t* h, const int64_t* idx64, const int32_t* idx32, const float* vals,
                            int32_t count, int32_t mode, float alpha, float eps, a0_stream_t stream_) {
  if (count == 0) return A0_OK;
  A0DeviceGuard guard(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (count <= K2P_MAX) return a0_launch_paths(h, idx64, idx32, vals, count, mode, alpha, eps, stream);
  const int chunk_log = h->D < K2R_MAXLOG ? h->D : K2R_MAXLOG;
  A0_LAUNCH(a0_k2b_write, 1, K2B_THREADS, 0, stream, 1, h->tree, h->P, chunk_log, h->N, idx64, idx32, vals, count, mode, alpha,
            eps, h->max_p, h->winner, h->dirty);
  A0_LAUNCH(a0_k2b_rebuild, (unsigned)(h->P >> chunk_log), K2R_THREADS, 0, stream, 1, h->tree, h->D, chunk_log, h->dirty,
            h->counter + A0_MAX_BATCHES);
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
