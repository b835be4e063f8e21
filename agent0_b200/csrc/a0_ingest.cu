// Host side of the replay shard: the ring index (which incoming frames are new, where they go,
// which records become sampleable, which must be evicted) and the staged ingest that executes an
// index plan on the device (one pinned H2D copy, then K2b marks and K1 append).
//
// Replaces the host bookkeeping of the reference's ReplayDataset.extend (agent0/deepq/replay.py:45-53:
// deque.extend + top + tail priorities) and the actor-side n-step tracker's "entry k-n+1 is emitted
// at step k" rule (agent0/deepq/agent.py:64-73).  The rules are those of agent0_b200/ring_index.py,
// which stays as the executable specification: tests/test_ring_index.py runs both implementations
// on the same streams and requires identical plans.
//
//   * sequence numbers are monotone int64; ring position = seq % capacity;
//   * a frame seq f is resident while f >= head_fs - NF;
//   * evict record q iff its slot is being reused or fs_at_append[q] - age_limit < head_fs - NF
//     (conservative and monotone, so the live records are the window [tail_q, head_q));
//   * with n-step gathering a record becomes sampleable when its (n-1)-th successor in the same
//     stream has been appended.
#include <algorithm>
#include <new>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "a0_common.cuh"

struct A0Stream {
  int32_t last_pos = -1;
  int64_t last_seq = -1;
  std::vector<int64_t> pending;   // appended, waiting for their (n-1)-th successor
  int64_t stack[A0_STACK] = {0, 0, 0, 0};
  bool has_stack = false;
};

struct a0_index {
  int64_t N, NF, n, age_limit;
  int64_t head_q = 0, tail_q = 0, head_fs = 0, top = 0;
  std::vector<int64_t> fs_at_append;
  std::vector<uint8_t> sampleable;
  std::unordered_map<int64_t, A0Stream> streams;
  std::unordered_set<int64_t> stale;
  // plan storage, valid until the next call
  std::vector<int32_t> new_frame_pos, rec_meta, marks;
  std::vector<int32_t> order;
  std::vector<int64_t> fs8, ready;

  int64_t max_chunk() const {
    const int64_t a = N / 2, b = (NF - 2 * age_limit) / (2 * A0_SLOTS);
    return std::max<int64_t>(1, std::min(a, b));
  }
};

extern "C" int a0_ix_create(a0_index_t** out, int64_t rec_capacity, int64_t frame_capacity, int32_t n_step,
                            int64_t age_limit) {
  A0_REQUIRE(out != nullptr, "a0_ix_create: out is NULL");
  A0_REQUIRE(rec_capacity >= 2 && frame_capacity >= 8, "a0_ix_create: capacities too small");
  A0_REQUIRE(n_step >= 1 && n_step <= A0_MAX_NSTEP, "a0_ix_create: n_step %d outside [1,%d]", n_step, A0_MAX_NSTEP);
  if (age_limit < 0) age_limit = std::max<int64_t>(8, std::min<int64_t>(32768, frame_capacity / 8));
  A0_REQUIRE(frame_capacity > 2 * age_limit, "a0_ix_create: frame ring (%lld) too small for the age limit (%lld)",
             (long long)frame_capacity, (long long)age_limit);
  a0_index* ix = new (std::nothrow) a0_index();
  if (!ix) { a0_set_error("a0_ix_create: out of host memory"); return A0_ENOMEM; }
  ix->N = rec_capacity; ix->NF = frame_capacity; ix->n = n_step; ix->age_limit = age_limit;
  ix->fs_at_append.assign((size_t)rec_capacity, 0);
  ix->sampleable.assign((size_t)rec_capacity, 0);
  *out = ix;
  return A0_OK;
}

extern "C" int a0_ix_destroy(a0_index_t* ix) {
  delete ix;
  return A0_OK;
}

extern "C" int a0_ix_state(a0_index_t* ix, int64_t* out) {
  A0_REQUIRE(ix && out, "a0_ix_state: NULL argument");
  out[A0_IX_HEAD_Q] = ix->head_q; out[A0_IX_TAIL_Q] = ix->tail_q; out[A0_IX_HEAD_FS] = ix->head_fs;
  out[A0_IX_TOP] = ix->top; out[A0_IX_REC_CAPACITY] = ix->N; out[A0_IX_FRAME_CAPACITY] = ix->NF;
  out[A0_IX_NSTEP] = ix->n; out[A0_IX_AGE_LIMIT] = ix->age_limit; out[A0_IX_MAX_CHUNK] = ix->max_chunk();
  return A0_OK;
}

extern "C" const uint8_t* a0_ix_sampleable(a0_index_t* ix) { return ix ? ix->sampleable.data() : nullptr; }

extern "C" int a0_ix_set_stack(a0_index_t* ix, int64_t stream, const int64_t* seq4) {
  A0_REQUIRE(ix && seq4, "a0_ix_set_stack: NULL argument");
  A0Stream& st = ix->streams[stream];
  for (int j = 0; j < A0_STACK; ++j) st.stack[j] = seq4[j];
  st.has_stack = true;
  return A0_OK;
}

extern "C" int a0_ix_get_stack(a0_index_t* ix, int64_t stream, int64_t* seq4) {
  A0_REQUIRE(ix && seq4, "a0_ix_get_stack: NULL argument");
  auto it = ix->streams.find(stream);
  if (it == ix->streams.end() || !it->second.has_stack) return 1;   // no stack yet (not an error)
  for (int j = 0; j < A0_STACK; ++j) seq4[j] = it->second.stack[j];
  return A0_OK;
}

// Groups of equal stream id in ascending id order, original order inside a group.
static void a0_group_by_stream(a0_index* ix, const int64_t* stream, int32_t m) {
  ix->order.resize((size_t)m);
  for (int32_t i = 0; i < m; ++i) ix->order[i] = i;
  std::stable_sort(ix->order.begin(), ix->order.end(), [&](int32_t a, int32_t b) { return stream[a] < stream[b]; });
}

extern "C" int a0_ix_resolve_shift(a0_index_t* ix, const int64_t* stream, const int64_t* n_new, int32_t m,
                                   int64_t* fs8_out) {
  A0_REQUIRE(ix && (m == 0 || (stream && n_new && fs8_out)), "a0_ix_resolve_shift: NULL argument");
  // new frames are numbered in (transition, frame) order from head_fs
  int64_t next = ix->head_fs;
  std::vector<int64_t>& off = ix->ready;     // scratch
  off.resize((size_t)m);
  for (int32_t i = 0; i < m; ++i) {
    A0_REQUIRE(n_new[i] >= 0 && n_new[i] <= A0_STACK, "a0_ix_resolve_shift: n_new[%d] = %lld outside [0,4]", i, (long long)n_new[i]);
    off[i] = next;
    next += n_new[i];
  }
  for (int32_t i = 0; i < m; ++i) {
    auto it = ix->streams.find(stream[i]);
    A0_REQUIRE(it != ix->streams.end() && it->second.has_stack,
               "a0_ix_resolve_shift: stream %lld has no observation stack yet (reset_streams first)", (long long)stream[i]);
  }
  // per stream, in commit order: obs stack = current stack, next stack = it shifted by the new frames
  for (int32_t i = 0; i < m; ++i) {
    A0Stream& st = ix->streams.find(stream[i])->second;
    int64_t* row = fs8_out + (size_t)i * A0_SLOTS;
    int64_t win[2 * A0_STACK];
    for (int j = 0; j < A0_STACK; ++j) { row[j] = st.stack[j]; win[j] = st.stack[j]; }
    const int c = (int)n_new[i];
    for (int j = 0; j < c; ++j) win[A0_STACK + j] = off[i] + j;
    for (int j = 0; j < A0_STACK; ++j) { row[A0_STACK + j] = win[c + j]; st.stack[j] = win[c + j]; }
  }
  return A0_OK;
}

extern "C" int a0_ix_plan(a0_index_t* ix, const int64_t* stream, const int64_t* fs8, int32_t m, int32_t n_new,
                          const int64_t* action, const double* reward, const uint8_t* done, a0_plan_t* out) {
  A0_REQUIRE(ix && out, "a0_ix_plan: NULL argument");
  A0_REQUIRE(m >= 0 && n_new >= 0, "a0_ix_plan: negative count");
  A0_REQUIRE(m == 0 || (stream && fs8 && action && reward && done), "a0_ix_plan: NULL array");
  A0_REQUIRE(m <= ix->max_chunk(), "a0_ix_plan: append of %d transitions is too large for the ring (max %lld): split it",
             m, (long long)ix->max_chunk());
  const int64_t N = ix->N, NF = ix->NF, n = ix->n;
  const int64_t q0 = ix->head_q;
  const int64_t new_head_fs = ix->head_fs + n_new, new_head_q = q0 + m;
  ix->marks.clear();
  // a record whose frames are already older than the age limit (an idle stream that resumes) can
  // never be a sampling start: the eviction bound below would not cover it
  for (int32_t i = 0; i < m; ++i) {
    int64_t mn = fs8[(size_t)i * A0_SLOTS];
    for (int j = 1; j < A0_SLOTS; ++j) mn = std::min(mn, fs8[(size_t)i * A0_SLOTS + j]);
    if (ix->head_fs - mn > ix->age_limit) ix->stale.insert(q0 + i);
  }
  // ---- evictions: a prefix of the live window ---------------------------------------------------
  const int64_t thr = new_head_fs - NF;
  while (ix->tail_q < q0) {
    const int64_t w = ix->tail_q, wp = w % N;
    if (!((w < new_head_q - N) || (ix->fs_at_append[wp] - ix->age_limit < thr))) break;
    ix->top -= ix->sampleable[wp];
    ix->sampleable[wp] = 0;
    ix->marks.push_back(~(int32_t)wp);
    ix->tail_q++;
  }
  // ---- links and sampleability, stream by stream ------------------------------------------------
  ix->rec_meta.resize((size_t)m * A0_REC_META_I32);
  int32_t* meta = ix->rec_meta.data();
  for (int32_t i = 0; i < m; ++i) {
    int32_t* mt = meta + (size_t)i * A0_REC_META_I32;
    mt[0] = (int32_t)((q0 + i) % N);
    mt[1] = -1;
    mt[2] = -1;
  }
  a0_group_by_stream(ix, stream, m);
  ix->ready.clear();
  bool any_newly = false;
  for (int32_t g0 = 0; g0 < m;) {
    int32_t g1 = g0;
    const int64_t sid = stream[ix->order[g0]];
    while (g1 < m && stream[ix->order[g1]] == sid) ++g1;
    A0Stream& st = ix->streams[sid];
    const int32_t first = ix->order[g0], last = ix->order[g1 - 1];
    if (st.last_seq >= std::max(ix->tail_q, q0 - N)) meta[(size_t)first * A0_REC_META_I32 + 1] = st.last_pos;
    for (int32_t g = g0; g + 1 < g1; ++g)
      meta[(size_t)ix->order[g] * A0_REC_META_I32 + 2] = meta[(size_t)ix->order[g + 1] * A0_REC_META_I32];
    if (n == 1) {
      for (int32_t g = g0; g < g1; ++g) ix->ready.push_back(q0 + ix->order[g]);
      any_newly = true;
    } else {
      for (int32_t g = g0; g < g1; ++g) st.pending.push_back(q0 + ix->order[g]);
      const int64_t nready = (int64_t)st.pending.size() - (n - 1);
      if (nready > 0) {
        ix->ready.insert(ix->ready.end(), st.pending.begin(), st.pending.begin() + nready);
        st.pending.erase(st.pending.begin(), st.pending.begin() + nready);
        any_newly = true;
      }
    }
    st.last_pos = meta[(size_t)last * A0_REC_META_I32];
    st.last_seq = q0 + last;
    g0 = g1;
  }
  for (int32_t i = 0; i < m; ++i) ix->fs_at_append[(size_t)((q0 + i) % N)] = ix->head_fs;
  ix->head_q = new_head_q;
  ix->head_fs = new_head_fs;
  if (any_newly) {
    const bool had_stale = !ix->stale.empty();
    std::unordered_set<int64_t> hit;
    for (int64_t r : ix->ready) {
      if (r < ix->tail_q) continue;                 // evicted before it ever became sampleable
      if (had_stale && ix->stale.count(r)) { hit.insert(r); continue; }
      const int64_t rp = r % N;
      ix->sampleable[rp] = 1;
      ix->top += 1;
      ix->marks.push_back((int32_t)rp);
    }
    if (had_stale) {
      for (auto it = ix->stale.begin(); it != ix->stale.end();) {
        if (*it < ix->tail_q || hit.count(*it)) it = ix->stale.erase(it);
        else ++it;
      }
    }
  }
  // ---- record metadata ----------------------------------------------------------------------------
  for (int32_t i = 0; i < m; ++i) {
    int32_t* mt = meta + (size_t)i * A0_REC_META_I32;
    mt[3] = (int32_t)(action[i] & 0x7fffffff) | (done[i] ? (int32_t)0x80000000 : 0);
    for (int j = 0; j < A0_SLOTS; ++j) {
      int64_t f = fs8[(size_t)i * A0_SLOTS + j] % NF;
      if (f < 0) f += NF;
      mt[4 + j] = (int32_t)f;
    }
    memcpy(mt + 12, reward + i, sizeof(double));
  }
  ix->new_frame_pos.resize((size_t)n_new);
  for (int32_t j = 0; j < n_new; ++j) ix->new_frame_pos[j] = (int32_t)((new_head_fs - n_new + j) % NF);
  out->m = m; out->n_new = n_new; out->n_marks = (int32_t)ix->marks.size();
  out->new_frame_pos = ix->new_frame_pos.data();
  out->rec_meta = ix->rec_meta.data();
  out->marks = ix->marks.data();
  return A0_OK;
}

// ------------------------------------------------------------------------------------------------
// content de-duplication for the reference-compatible ingest
// ------------------------------------------------------------------------------------------------
// The reference's entries are self-contained 8-frame blobs concat(st, st_next) (agent.py:78-81), 7 of
// which normally repeat frames of the same env's previous entry.  A frame is matched -- by a 64-bit
// hash first, then by full byte comparison, so a match is always bit-exact -- against the 8 frames of
// the stream's previous entry and the earlier frames of its own entry; unmatched frames are new.
// Specification: ring_index.ContentDeduper (numpy, 260 ms per 1280-transition extend); this is the
// same rule at memory speed (tests/test_ring_index.py drives both and requires identical output).
struct A0Tail {
  std::vector<uint8_t> frames;      // owned copy of the last entry's 8 frames (filled when a call ends)
  const uint8_t* cur = nullptr;     // the last entry's frames during a call (caller's buffer or `frames`)
  int64_t seq[A0_SLOTS];
  uint64_t hash[A0_SLOTS];
  bool valid = false;
};
struct a0_dedupe {
  int32_t F;
  std::unordered_map<int64_t, A0Tail> tail;
};

static inline uint64_t a0_frame_hash(const uint8_t* p, int32_t F) {
  // four independent multiply-rotate lanes over 64-bit words (frames are multiples of 16 bytes)
  uint64_t h0 = 0x9E3779B97F4A7C15ull, h1 = 0xC2B2AE3D27D4EB4Full, h2 = 0x165667B19E3779F9ull, h3 = 0x27D4EB2F165667C5ull;
  const int n = F / 8;
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    uint64_t w0, w1, w2, w3;
    memcpy(&w0, p + 8 * i, 8); memcpy(&w1, p + 8 * i + 8, 8); memcpy(&w2, p + 8 * i + 16, 8); memcpy(&w3, p + 8 * i + 24, 8);
    h0 = (h0 ^ w0) * 0x9FB21C651E98DF25ull; h0 = (h0 << 29) | (h0 >> 35);
    h1 = (h1 ^ w1) * 0x9FB21C651E98DF25ull; h1 = (h1 << 29) | (h1 >> 35);
    h2 = (h2 ^ w2) * 0x9FB21C651E98DF25ull; h2 = (h2 << 29) | (h2 >> 35);
    h3 = (h3 ^ w3) * 0x9FB21C651E98DF25ull; h3 = (h3 << 29) | (h3 >> 35);
  }
  for (; i < n; ++i) {
    uint64_t w;
    memcpy(&w, p + 8 * i, 8);
    h0 = (h0 ^ w) * 0x9FB21C651E98DF25ull; h0 = (h0 << 29) | (h0 >> 35);
  }
  uint64_t h = h0 ^ (h1 * 3) ^ (h2 * 5) ^ (h3 * 7);
  h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
  return h;
}

extern "C" int a0_dd_create(a0_dedupe_t** out, int32_t frame_bytes) {
  A0_REQUIRE(out != nullptr, "a0_dd_create: out is NULL");
  A0_REQUIRE(frame_bytes > 0 && frame_bytes % 8 == 0, "a0_dd_create: frame_bytes %d must be a positive multiple of 8", frame_bytes);
  a0_dedupe* dd = new (std::nothrow) a0_dedupe();
  if (!dd) { a0_set_error("a0_dd_create: out of host memory"); return A0_ENOMEM; }
  dd->F = frame_bytes;
  *out = dd;
  return A0_OK;
}

extern "C" int a0_dd_destroy(a0_dedupe_t* dd) {
  delete dd;
  return A0_OK;
}

extern "C" int a0_dd_resolve(a0_dedupe_t* dd, a0_index_t* ix, const int64_t* stream, const uint8_t* frames, int32_t m,
                             int64_t* fs8_out, int64_t* new_src_out, int32_t* n_new_out) {
  A0_REQUIRE(dd && ix && n_new_out, "a0_dd_resolve: NULL handle");
  A0_REQUIRE(m >= 0, "a0_dd_resolve: negative count");
  A0_REQUIRE(m == 0 || (stream && frames && fs8_out && new_src_out), "a0_dd_resolve: NULL array");
  const size_t F = (size_t)dd->F;
  int64_t next_fs = ix->head_fs;
  const int64_t resident_from = ix->head_fs - ix->NF;     // frame_resident(f): f >= head_fs - NF
  int32_t n_new = 0;
  std::vector<int64_t> touched;
  for (int32_t t = 0; t < m; ++t) {
    A0Tail& tl = dd->tail[stream[t]];
    const uint8_t* ent = frames + (size_t)t * A0_SLOTS * F;
    int64_t* row = fs8_out + (size_t)t * A0_SLOTS;
    uint64_t hs[A0_SLOTS];
    for (int j = 0; j < A0_SLOTS; ++j) hs[j] = a0_frame_hash(ent + j * F, dd->F);
    const uint8_t* prev = tl.valid ? (tl.cur ? tl.cur : tl.frames.data()) : nullptr;
    for (int j = 0; j < A0_SLOTS; ++j) {
      int64_t hit = -1;
      if (prev) {
        for (int c = 0; c < A0_SLOTS && hit < 0; ++c)
          if (tl.hash[c] == hs[j] && next_fs - tl.seq[c] < ix->age_limit && tl.seq[c] >= resident_from &&
              memcmp(prev + c * F, ent + j * F, F) == 0)
            hit = tl.seq[c];
      }
      for (int c = 0; c < j && hit < 0; ++c)
        if (hs[c] == hs[j] && memcmp(ent + c * F, ent + j * F, F) == 0) hit = row[c];
      if (hit < 0) {
        hit = next_fs++;
        new_src_out[n_new++] = (int64_t)t * A0_SLOTS + j;
      }
      row[j] = hit;
    }
    if (!tl.valid || tl.cur == nullptr) touched.push_back(stream[t]);
    tl.cur = ent;
    tl.valid = true;
    for (int j = 0; j < A0_SLOTS; ++j) { tl.seq[j] = row[j]; tl.hash[j] = hs[j]; }
  }
  // private copies of the stream tails: the caller may reuse its buffer after the call
  for (int64_t sid : touched) {
    A0Tail& tl = dd->tail[sid];
    if (tl.cur) {
      tl.frames.assign(tl.cur, tl.cur + A0_SLOTS * F);
      tl.cur = nullptr;
    }
  }
  *n_new_out = n_new;
  return A0_OK;
}

// ------------------------------------------------------------------------------------------------
// staged execution of a plan on the device
// ------------------------------------------------------------------------------------------------
static inline size_t a0_up256(size_t x) { return (x + 255) & ~(size_t)255; }

static int a0_stage_reserve(a0_replay* h, int turn, size_t bytes) {
  A0Staging& s = h->staging[turn];
  if (s.event) A0_CUDA(cudaEventSynchronize(s.event));     // the previous use of this buffer has been consumed
  if (s.capacity >= bytes) return A0_OK;
  const size_t cap = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 20);
  if (s.host) { cudaFreeHost(s.host); s.host = nullptr; }
  if (s.dev) { cudaFree(s.dev); s.dev = nullptr; }
  s.capacity = 0;
  A0_CUDA(cudaMallocHost((void**)&s.host, cap));
  A0_CUDA(cudaMalloc((void**)&s.dev, cap));
  if (!s.event) A0_CUDA(cudaEventCreateWithFlags(&s.event, cudaEventDisableTiming));
  if (!s.copied) A0_CUDA(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
  s.capacity = cap;
  return A0_OK;
}

// dyn.dyn != NULL: the launch that appends also publishes {top, beta, sum_offset} for the sampler
// (what a0_rb_set_dynamic would otherwise launch a kernel for).
int a0_ingest_plan_impl(a0_replay_t* h, const a0_plan_t* plan, const uint8_t* frames,
                        const int64_t* new_frame_src, int32_t flags, float alpha, a0_stream_t stream_,
                        const A0Dyn& dyn) {
  A0_REQUIRE(h && plan, "a0_rb_ingest_plan: NULL argument");
  const int32_t n_new = plan->n_new, m = plan->m, k = plan->n_marks;
  if (n_new == 0 && m == 0 && k == 0) {
    if (!dyn.dyn) return A0_OK;
    A0DeviceGuard guard(h->device);
    return a0_append_launch(h, nullptr, nullptr, nullptr, 0, nullptr, 0, dyn, (cudaStream_t)stream_);
  }
  A0_REQUIRE(n_new == 0 || frames, "a0_rb_ingest_plan: frames is NULL");
  const bool on_device = flags & A0_INGEST_FRAMES_ON_DEVICE;
  const bool pinned = flags & A0_INGEST_FRAMES_PINNED;
  A0_REQUIRE(!pinned || new_frame_src == nullptr, "a0_rb_ingest_plan: pinned frames must already be in allocation order");
  // device frames + new_frame_src: the picks are uploaded as int32 and K1 reads frames[src[j]] (a0_ex_extend)
  const bool dev_pick = on_device && new_frame_src != nullptr && n_new > 0;
  A0DeviceGuard guard(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t F = (size_t)h->F;
  // successor stride of this append (rec_meta: {pos, link_from, ...}): the same for every record when
  // the actors step in lockstep; K3 uses it as a (verified) guess, see a0_spec_fetch
  if (m > 0) {
    int64_t stride = -1;
    for (int32_t i = 0; i < m && stride != 0; ++i) {
      const int32_t* mt = plan->rec_meta + (size_t)i * A0_REC_META_I32;
      if (mt[1] < 0) continue;                         // first record of its stream
      int64_t dlt = (int64_t)mt[0] - mt[1];
      if (dlt <= 0) dlt += h->N;
      stride = (stride < 0 || stride == dlt) ? dlt : 0;
    }
    if (stride >= 0) h->stride_hint = stride;
  }
  const bool stage_frames = n_new > 0 && !on_device && !pinned;
  size_t sec[6];
  sec[0] = 0;
  sec[1] = sec[0] + a0_up256(on_device ? 0 : (size_t)n_new * F);
  sec[2] = sec[1] + a0_up256((size_t)n_new * 4);
  sec[3] = sec[2] + a0_up256((size_t)m * A0_REC_META_I32 * 4);
  sec[5] = sec[3] + a0_up256((size_t)k * 4);
  sec[4] = sec[5] + a0_up256(dev_pick ? (size_t)n_new * 4 : 0);     // sec[4] stays "end of the upload"; [sec[5], sec[4]) = picks
  const int turn = h->staging_turn;
  h->staging_turn ^= 1;
  int rc = a0_stage_reserve(h, turn, sec[4]);
  if (rc) return rc;
  A0Staging& s = h->staging[turn];
  if (stage_frames) {
    for (int32_t j = 0; j < n_new; ++j) {
      const size_t src = new_frame_src ? (size_t)new_frame_src[j] : (size_t)j;
      memcpy(s.host + sec[0] + (size_t)j * F, frames + src * F, F);
    }
  }
  if (n_new) memcpy(s.host + sec[1], plan->new_frame_pos, (size_t)n_new * 4);
  if (m) memcpy(s.host + sec[2], plan->rec_meta, (size_t)m * A0_REC_META_I32 * 4);
  if (k) memcpy(s.host + sec[3], plan->marks, (size_t)k * 4);
  if (dev_pick) {
    int32_t* pick = reinterpret_cast<int32_t*>(s.host + sec[5]);
    for (int32_t j = 0; j < n_new; ++j) pick[j] = (int32_t)new_frame_src[j];
  }
  // The H2D copies may run on the shard's own copy stream, so that they overlap whatever `stream`
  // is still executing (the previous step's kernels); `stream` then waits for them.  The staging
  // buffer they fill was released by a0_stage_reserve (its last consumer has finished).
  cudaStream_t cs = stream;
  if (flags & A0_INGEST_COPY_STREAM) {
    if (!h->copy_stream) A0_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    cs = h->copy_stream;
  }
  if (stage_frames) {
    A0_CUDA(cudaMemcpyAsync(s.dev, s.host, sec[4], cudaMemcpyHostToDevice, cs));
  } else {
    if (pinned && n_new) A0_CUDA(cudaMemcpyAsync(s.dev + sec[0], frames, (size_t)n_new * F, cudaMemcpyHostToDevice, cs));
    A0_CUDA(cudaMemcpyAsync(s.dev + sec[1], s.host + sec[1], sec[4] - sec[1], cudaMemcpyHostToDevice, cs));
  }
  if (cs != stream) {
    A0_CUDA(cudaEventRecord(s.copied, cs));
    A0_CUDA(cudaStreamWaitEvent(stream, s.copied, 0));
  }
  const uint8_t* dev_frames = n_new ? (on_device ? frames : s.dev + sec[0]) : nullptr;
  const int32_t* dev_pos = n_new ? (const int32_t*)(s.dev + sec[1]) : nullptr;
  const int32_t* dev_meta = m ? (const int32_t*)(s.dev + sec[2]) : nullptr;
  const int32_t* dev_pick_idx = dev_pick ? (const int32_t*)(s.dev + sec[5]) : nullptr;
  A0_REQUIRE(((uintptr_t)dev_frames & 15) == 0, "a0_rb_ingest_plan: device frames must be 16-byte aligned");
  // A step-sized append (a few hundred marks, tens of frames) goes out as ONE launch: marks and
  // append touch disjoint state.  Bulk fills keep the two launches (cluster marks, 128-thread K1 CTAs).
  rc = (k && a0_option_fused_ingest()) ? a0_launch_mark_append(h, (const int32_t*)(s.dev + sec[3]), k, alpha, dev_frames, dev_pos, dev_pick_idx, n_new,
                                                               dev_meta, m, dyn, stream)
                                       : A0_NOFIT;
  if (rc == A0_NOFIT) {
    if (k) {
      rc = a0_pt_mark(h, (const int32_t*)(s.dev + sec[3]), k, alpha, stream_);
      if (rc) return rc;
    }
    rc = a0_append_launch(h, dev_frames, dev_pos, dev_pick_idx, n_new, dev_meta, m, dyn, stream);
  }
  if (rc) return rc;
  A0_CUDA(cudaEventRecord(s.event, stream));
  return A0_OK;
}

extern "C" int a0_rb_ingest_plan(a0_replay_t* h, const a0_plan_t* plan, const uint8_t* frames,
                                 const int64_t* new_frame_src, int32_t flags, float alpha, a0_stream_t stream) {
  const A0Dyn none = {nullptr, 0.0f, 0.0f, 0.0f};
  return a0_ingest_plan_impl(h, plan, frames, new_frame_src, flags, alpha, stream, none);
}

// beta == NULL: nothing is published.  Otherwise the LAST chunk's append publishes
// {ix->top after the append, *beta, compat_sum ? N - top : 0}.
static int a0_ingest_steps_impl(a0_replay_t* h, a0_index_t* ix, const int64_t* stream, const int64_t* n_new,
                                const uint8_t* new_frames, int32_t flags, const int64_t* action, const double* reward,
                                const uint8_t* done, int32_t m, float alpha, const float* beta, int32_t compat_sum,
                                a0_stream_t cuda_stream, const char* who) {
  A0_REQUIRE(h && ix, "%s: NULL handle", who);
  A0_REQUIRE(m >= 0, "%s: negative count", who);
  A0_REQUIRE(ix->N == h->N && ix->NF == h->NF, "%s: index and shard capacities differ", who);
  { int frc = a0_check_fault(h, who); if (frc) return frc; }
  const size_t F = (size_t)h->F;
  const int64_t step = ix->max_chunk();
  size_t frame_off = 0;
  A0Dyn dyn = {nullptr, 0.0f, 0.0f, 0.0f};
  for (int64_t lo = 0; lo < m || (lo == 0 && beta); lo += step) {
    const int32_t cnt = (int32_t)std::max<int64_t>(0, std::min<int64_t>(step, m - lo));
    ix->fs8.resize((size_t)cnt * A0_SLOTS);
    int rc = a0_ix_resolve_shift(ix, stream + lo, n_new + lo, cnt, ix->fs8.data());
    if (rc) return rc;
    int64_t total_new = 0;
    for (int32_t i = 0; i < cnt; ++i) total_new += n_new[lo + i];
    a0_plan_t plan;
    rc = a0_ix_plan(ix, stream + lo, ix->fs8.data(), cnt, (int32_t)total_new, action + lo, reward + lo, done + lo, &plan);
    if (rc) return rc;
    if (beta && lo + step >= m) {
      dyn.dyn = h->dyn;
      dyn.top = (float)ix->top;
      dyn.beta = *beta;
      dyn.sum_offset = compat_sum ? (float)(ix->N - ix->top) : 0.0f;
    }
    rc = a0_ingest_plan_impl(h, &plan, new_frames ? new_frames + frame_off * F : nullptr, nullptr, flags, alpha, cuda_stream,
                             dyn);
    if (rc) return rc;
    frame_off += (size_t)total_new;
    if (m == 0) break;
  }
  return A0_OK;
}

extern "C" int a0_rb_ingest_steps(a0_replay_t* h, a0_index_t* ix, const int64_t* stream, const int64_t* n_new,
                                  const uint8_t* new_frames, int32_t flags, const int64_t* action,
                                  const double* reward, const uint8_t* done, int32_t m, float alpha,
                                  a0_stream_t cuda_stream) {
  return a0_ingest_steps_impl(h, ix, stream, n_new, new_frames, flags, action, reward, done, m, alpha, nullptr, 0,
                              cuda_stream, "a0_rb_ingest_steps");
}

extern "C" int a0_rb_ingest_steps_dyn(a0_replay_t* h, a0_index_t* ix, const int64_t* stream, const int64_t* n_new,
                                      const uint8_t* new_frames, int32_t flags, const int64_t* action,
                                      const double* reward, const uint8_t* done, int32_t m, float alpha, float beta,
                                      int32_t compat_sum, a0_stream_t cuda_stream) {
  return a0_ingest_steps_impl(h, ix, stream, n_new, new_frames, flags, action, reward, done, m, alpha, &beta, compat_sum,
                              cuda_stream, "a0_rb_ingest_steps_dyn");
}
