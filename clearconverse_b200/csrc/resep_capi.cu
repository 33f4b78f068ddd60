// C ABI of libresep_b200.so (include/resep_b200.h): handle lifetime, weight upload, per-shape
// plans, workspace carving and the forward-pass orchestration that restates
// SepformerSeparation.separate_batch (speechbrain/inference/separation.py), called by the
// reference at /root/reference/back/api.py:1077.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>

#include "resep_internal.cuh"
#include "resep_tc.cuh"

namespace resep {

static thread_local std::string g_create_err;

int set_err(ResepHandle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------ weights
struct ArenaBuilder {
  std::vector<unsigned char> host;
  size_t add(const void* src, size_t bytes) {
    size_t off = align_up(host.size(), 256);
    host.resize(off + bytes);
    std::memcpy(host.data() + off, src, bytes);
    return off;
  }
  size_t add_f32(const float* src, size_t n) { return add(src, n * sizeof(float)); }
  size_t add_tf32(const float* src, size_t n) {   // round to nearest, ties away (== cvt.rna.tf32.f32)
    std::vector<float> tmp(n);
    for (size_t i = 0; i < n; ++i) {
      uint32_t u;
      std::memcpy(&u, &src[i], 4);
      if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
      std::memcpy(&tmp[i], &u, 4);
    }
    return add(tmp.data(), n * sizeof(float));
  }
  static float tf32_rna(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
    std::memcpy(&x, &u, 4);
    return x;
  }
  size_t add_tf32_lo(const float* src, size_t n) {   // tf32(W - tf32(W)): second term of the split weights
    std::vector<float> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = tf32_rna(src[i] - tf32_rna(src[i]));
    return add(tmp.data(), n * sizeof(float));
  }
  size_t add_bf16_lo(const float* src, size_t n) {
    std::vector<bf16> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = __float2bfloat16_rn(src[i] - __bfloat162float(__float2bfloat16_rn(src[i])));
    return add(tmp.data(), n * sizeof(bf16));
  }
  size_t add_f16(const float* src, size_t n, bool lo) {   // IEEE fp16: fp16(W), or fp16(W - fp16(W))
    std::vector<__half> tmp(n);
    for (size_t i = 0; i < n; ++i) {
      const __half hi = __float2half_rn(src[i]);
      tmp[i] = lo ? __float2half_rn(src[i] - __half2float(hi)) : hi;
    }
    return add(tmp.data(), n * sizeof(__half));
  }
  size_t add_bf16(const float* src, size_t n) {
    std::vector<bf16> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = __float2bfloat16_rn(src[i]);
    return add(tmp.data(), n * sizeof(bf16));
  }
};

static int upload_weights(ResepHandle* h, const ResepWeights* w) {
  if (!w || !w->enc_w || !w->dec_w || !w->prelu_a || !w->fc_w || !w->fc_b || !w->pe)
    return set_err(h, RESEP_EINVAL, "null weight pointer");
  if (w->pe_rows < CHUNK) return set_err(h, RESEP_EINVAL, "pe_rows must be >= 150");
  ArenaBuilder ab;
  struct Fix { const void** slot; size_t off; };
  std::vector<Fix> fixes;
  WeightsDev& d = h->w;
  auto F = [&](const float*& slot, const float* src, size_t n) {
    fixes.push_back({reinterpret_cast<const void**>(&slot), ab.add_f32(src, n)});
  };
  auto Bf = [&](const bf16*& slot, const bf16*& slot_lo, const float* src, size_t n) {
    fixes.push_back({reinterpret_cast<const void**>(&slot), ab.add_bf16(src, n)});
    fixes.push_back({reinterpret_cast<const void**>(&slot_lo), ab.add_bf16_lo(src, n)});
  };
  auto Hf = [&](const bf16* (&slot)[2], const float* src, size_t n) {
    fixes.push_back({reinterpret_cast<const void**>(&slot[0]), ab.add_f16(src, n, false)});
    fixes.push_back({reinterpret_cast<const void**>(&slot[1]), ab.add_f16(src, n, true)});
  };
  auto Tf = [&](const float*& slot, const float*& slot_lo, const float* src, size_t n) {
    fixes.push_back({reinterpret_cast<const void**>(&slot), ab.add_tf32(src, n)});
    fixes.push_back({reinterpret_cast<const void**>(&slot_lo), ab.add_tf32_lo(src, n)});
  };
  F(d.enc_w, w->enc_w, D * KSZ);
  F(d.dec_w, w->dec_w, D * KSZ);
  F(d.prelu_a, w->prelu_a, 1);
  F(d.fc_w, w->fc_w, NSPK * D * D);
  F(d.fc_b, w->fc_b, NSPK * D);
  F(d.pe, w->pe, (size_t)w->pe_rows * D);
  Bf(d.fc_w_bf, d.fc_w_bl, w->fc_w, NSPK * D * D);
  Hf(d.fc_w_h, w->fc_w, NSPK * D * D);
  Tf(d.fc_w_tf, d.fc_w_lo, w->fc_w, NSPK * D * D);
  d.pe_rows = w->pe_rows;
  const ResepBlockWeights* src_blocks[3] = {&w->seg[0], &w->seg[1], &w->mem[0]};
  constexpr size_t POST_PAR = 4 * D + FFN + 3 * D + 2 * D + FFN;
  h->host_par.assign(3 * NL * POST_PAR, 0.f);
  for (int b = 0; b < 3; ++b) {
    const ResepBlockWeights& sb = *src_blocks[b];
    BlockDev& db = d.blk[b];
    if (!sb.final_norm_w || !sb.final_norm_b || !sb.gln_w || !sb.gln_b) return set_err(h, RESEP_EINVAL, "null block weight");
    F(db.fn_w, sb.final_norm_w, D); F(db.fn_b, sb.final_norm_b, D);
    F(db.gln_w, sb.gln_w, D); F(db.gln_b, sb.gln_b, D);
    for (int l = 0; l < NL; ++l) {
      const ResepLayerWeights& s = sb.layers[l];
      LayerDev& t = db.layers[l];
      if (!s.norm1_w || !s.norm1_b || !s.in_proj_w || !s.in_proj_b || !s.out_proj_w || !s.out_proj_b || !s.norm2_w ||
          !s.norm2_b || !s.ffn1_w || !s.ffn1_b || !s.ffn2_w || !s.ffn2_b)
        return set_err(h, RESEP_EINVAL, "null layer weight");
      {
        float* hp = h->host_par.data() + (size_t)(b * NL + l) * POST_PAR;
        std::memcpy(hp, s.out_proj_b, D * 4); std::memcpy(hp + D, s.norm2_w, D * 4); std::memcpy(hp + 2 * D, s.norm2_b, D * 4);
        std::memcpy(hp + 3 * D, s.ffn2_b, D * 4); std::memcpy(hp + 4 * D, s.ffn1_b, FFN * 4);
        std::memcpy(hp + 7 * D + FFN, s.norm1_w, D * 4); std::memcpy(hp + 8 * D + FFN, s.norm1_b, D * 4);
        t.h_post_par = hp;
        t.h_in_b = hp + 4 * D + FFN;
      }
      F(t.norm1_w, s.norm1_w, D); F(t.norm1_b, s.norm1_b, D);
      F(t.in_w, s.in_proj_w, 3 * D * D); F(t.in_b, s.in_proj_b, 3 * D);
      F(t.out_w, s.out_proj_w, D * D); F(t.out_b, s.out_proj_b, D);
      F(t.norm2_w, s.norm2_w, D); F(t.norm2_b, s.norm2_b, D);
      F(t.f1_w, s.ffn1_w, (size_t)FFN * D); F(t.f1_b, s.ffn1_b, FFN);
      F(t.f2_w, s.ffn2_w, (size_t)D * FFN); F(t.f2_b, s.ffn2_b, D);
      {  // the bf16 path keeps q, k, v of a head adjacent in the qkv buffer ([head][q|k|v][16], see QKV_HEAD_STRIDE;
         // k pre-scaled by QK_PRESCALE):
         // a row permutation of the in-projection, so that one TMA box per (chunk, head) fetches all three slices
        std::vector<float> pw((size_t)3 * D * D), pb(3 * D);
        for (int part = 0; part < 3; ++part)
          for (int hd = 0; hd < NH; ++hd)
            for (int e = 0; e < DH; ++e) {
              const int r_old = part * D + hd * DH + e, r_new = hd * QKV_HEAD_STRIDE + part * DH + e;
              const float sc = part == 1 ? QK_PRESCALE : 1.f;
              for (int c = 0; c < D; ++c) pw[(size_t)r_new * D + c] = s.in_proj_w[(size_t)r_old * D + c] * sc;
              pb[r_new] = s.in_proj_b[r_old] * sc;
            }
        Bf(t.in_w_bf, t.in_w_bl, pw.data(), 3 * D * D);
        Hf(t.in_w_h, pw.data(), 3 * D * D);
        F(t.in_b_hi, pb.data(), 3 * D);
        std::memcpy(h->host_par.data() + (size_t)(b * NL + l) * POST_PAR + 4 * D + FFN, pb.data(), 3 * D * 4);
      }
      Bf(t.out_w_bf, t.out_w_bl, s.out_proj_w, D * D);
      Bf(t.f1_w_bf, t.f1_w_bl, s.ffn1_w, (size_t)FFN * D);
      Bf(t.f2_w_bf, t.f2_w_bl, s.ffn2_w, (size_t)D * FFN);
      Hf(t.out_w_h, s.out_proj_w, D * D);
      Hf(t.f1_w_h, s.ffn1_w, (size_t)FFN * D);
      {  // norm2 folded into FFN1 for k_post2_tc (see LayerDev::f1g_w_bf)
        std::vector<float> wg((size_t)FFN * D);
        float* b1g = h->host_par.data() + (size_t)(b * NL + l) * POST_PAR + 9 * D + FFN;
        for (int hh = 0; hh < FFN; ++hh) {
          double acc = s.ffn1_b[hh];
          for (int k = 0; k < D; ++k) {
            wg[(size_t)hh * D + k] = s.ffn1_w[(size_t)hh * D + k] * s.norm2_w[k];
            acc += (double)s.ffn1_w[(size_t)hh * D + k] * s.norm2_b[k];
          }
          b1g[hh] = (float)acc;
        }
        t.h_b1g = b1g;
        std::vector<float> pp(2 * D + FFN);
        std::memcpy(pp.data(), s.out_proj_b, D * 4); std::memcpy(pp.data() + D, s.ffn2_b, D * 4); std::memcpy(pp.data() + 2 * D, b1g, FFN * 4);
        F(t.post_par, pp.data(), pp.size());
        Bf(t.f1g_w_bf, t.f1g_w_bl, wg.data(), (size_t)FFN * D);
        Hf(t.f1g_w_h, wg.data(), (size_t)FFN * D);
      }
      Hf(t.f2_w_h, s.ffn2_w, (size_t)D * FFN);
      Tf(t.in_w_tf, t.in_w_lo, s.in_proj_w, 3 * D * D);
      Tf(t.out_w_tf, t.out_w_lo, s.out_proj_w, D * D);
      Tf(t.f1_w_tf, t.f1_w_lo, s.ffn1_w, (size_t)FFN * D);
      Tf(t.f2_w_tf, t.f2_w_lo, s.ffn2_w, (size_t)D * FFN);
    }
  }
  if (h->arena == nullptr || h->arena_bytes < ab.host.size()) {
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr;
    RESEP_CUDA(h, cudaMalloc(&h->arena, ab.host.size()));
    h->arena_bytes = ab.host.size();
  }
  RESEP_CUDA(h, cudaMemcpy(h->arena, ab.host.data(), ab.host.size(), cudaMemcpyHostToDevice));
  for (auto& f : fixes) *f.slot = static_cast<unsigned char*>(h->arena) + f.off;
  return RESEP_OK;
}

// ------------------------------------------------------------------------------ plans
static void free_plan(Plan* p) {   // handle teardown only: releases the blocks for real
  if (!p) return;
  if (p->dev) cudaFree(p->dev);
  if (p->host) cudaFreeHost(p->host);
  delete p;
}

// Eviction: the plan's blocks go back to the handle's pool.  Work already enqueued that reads the device tables is
// safe: a block is only rewritten by a later plan's upload, which is stream-ordered behind that work on the same
// stream, and a handle is thread-compatible (one caller at a time).  Lanes on OTHER streams keep their plans alive by
// using them (LRU), and the table is large enough (64) that a plan evicted here has not been launched for 63 shapes.
static void recycle_plan(ResepHandle* h, Plan* p) {
  if (!p) return;
  if (p->dev && p->host) h->plan_pool.push_back({p->dev, p->host, p->cap_bytes});
  else { if (p->dev) cudaFree(p->dev); if (p->host) cudaFreeHost(p->host); }
  delete p;
}

static void drop_graphs_of(ResepHandle* h, const Plan* plan = nullptr) {   // plan == nullptr: every graph
  for (size_t i = 0; i < h->graphs.size();) {
    if (plan == nullptr || h->graphs[i].plan == plan || h->graphs[i].plan == nullptr) {
      if (h->graphs[i].exec) cudaGraphExecDestroy(h->graphs[i].exec);
      h->graphs.erase(h->graphs.begin() + i);
    } else ++i;
  }
}

static int get_plan(ResepHandle* h, int B, const int64_t* item_off, const int64_t* item_len, int batch_mode,
                    cudaStream_t st, Plan** out) {
  std::vector<int64_t> key;
  key.reserve(2 + 2 * (size_t)B);
  key.push_back(B);
  key.push_back(batch_mode);
  for (int i = 0; i < B; ++i) { key.push_back(item_off[i]); key.push_back(item_len[i]); }
  h->tick++;
  for (Plan* p : h->plans)
    if (p->key == key) { p->last_use = h->tick; *out = p; return RESEP_OK; }

  Plan* p = new (std::nothrow) Plan();
  if (!p) return set_err(h, RESEP_EINVAL, "out of host memory");
  p->key = key;
  p->B = B;
  std::vector<int> item_L(B), item_row0(B), item_S(B);
  int64_t chunks = 0;
  for (int i = 0; i < B; ++i) {
    const int64_t T = item_len[i];
    if (T < KSZ) {
      delete p;
      return set_err(h, RESEP_ESHORT, "item " + std::to_string(i) + " has " + std::to_string(T) +
                                          " samples; kernel size (16) can't be greater than actual input size");
    }
    const int64_t L = (T - KSZ) / STRIDE + 1;
    if (L > 2000000000LL / D) { delete p; return set_err(h, RESEP_EINVAL, "item too long"); }
    item_L[i] = (int)L;
    item_S[i] = (int)(L / CHUNK + 1);   // rest = K - L % K is in [1, K]: a full zero chunk when L % K == 0
    if (batch_mode == RESEP_BATCH_SPAN_EXACT && L % CHUNK == 0) item_S[i] = (int)(L / CHUNK);   // inner span of a longer recording
    item_row0[i] = (int)(chunks * CHUNK);
    chunks += item_S[i];
    if (chunks * CHUNK > 2000000000LL) { delete p; return set_err(h, RESEP_EINVAL, "batch too large (token rows overflow int32)"); }
  }
  p->n_chunks = chunks;
  p->M = chunks * CHUNK;
  p->maskdec_ok = true;
  for (int i = 0; i < B; ++i) {
    if (8LL * CHUNK * item_S[i] < item_len[i] || item_off[i] + item_len[i] >= (1LL << 30)) p->maskdec_ok = false;
  }
  std::vector<int> chunk_item(chunks), chunk_frame0(chunks), mem_pos(chunks);
  std::vector<int> mem_seq_off;
  {
    int64_t c = 0;
    for (int i = 0; i < B; ++i)
      for (int s = 0; s < item_S[i]; ++s, ++c) { chunk_item[c] = i; chunk_frame0[c] = s * CHUNK; }
  }
  if (batch_mode == RESEP_BATCH_COUPLED) {
    mem_seq_off = {0, (int)chunks};
  } else {
    mem_seq_off.push_back(0);
    for (int i = 0; i < B; ++i) mem_seq_off.push_back(mem_seq_off.back() + item_S[i]);
  }
  p->n_mem_seq = (int)mem_seq_off.size() - 1;
  std::vector<int> mem_tile_seq, mem_tile_q0;
  for (int s = 0; s < p->n_mem_seq; ++s) {
    const int len = mem_seq_off[s + 1] - mem_seq_off[s];
    p->max_mem_len = std::max(p->max_mem_len, len);
    for (int j = 0; j < len; ++j) mem_pos[mem_seq_off[s] + j] = j;
    for (int q0 = 0; q0 < len; q0 += 128) { mem_tile_seq.push_back(s); mem_tile_q0.push_back(q0); }
  }
  p->n_mem_tiles = (int)mem_tile_seq.size();
  if (p->max_mem_len > h->w.pe_rows) {
    delete p;
    return set_err(h, RESEP_EPOS, "memory sequence of " + std::to_string(p->max_mem_len) +
                                      " chunks exceeds the positional-encoding table");
  }
  std::vector<int> dec_tile_item, dec_tile_slot0;
  for (int i = 0; i < B; ++i) {
    const int64_t slots = (item_len[i] + STRIDE - 1) / STRIDE;
    for (int64_t q = 0; q < slots; q += 31) { dec_tile_item.push_back(i); dec_tile_slot0.push_back((int)q); }   // DEC_SLOTS of k_decoder
  }
  p->n_dec_tiles = (int)dec_tile_item.size();

  // pack every table into one host blob -> one device allocation, one copy
  std::vector<unsigned char> blob;
  auto put = [&](const void* src, size_t bytes) {
    size_t off = align_up(blob.size(), 16);
    blob.resize(off + bytes);
    if (bytes) std::memcpy(blob.data() + off, src, bytes);
    return off;
  };
  const size_t o_off = put(item_off, sizeof(int64_t) * B), o_len = put(item_len, sizeof(int64_t) * B);
  const size_t o_L = put(item_L.data(), sizeof(int) * B), o_row0 = put(item_row0.data(), sizeof(int) * B);
  const size_t o_ci = put(chunk_item.data(), sizeof(int) * chunks), o_cf = put(chunk_frame0.data(), sizeof(int) * chunks);
  const size_t o_mp = put(mem_pos.data(), sizeof(int) * chunks);
  const size_t o_ms = put(mem_seq_off.data(), sizeof(int) * mem_seq_off.size());
  const size_t o_mts = put(mem_tile_seq.data(), sizeof(int) * mem_tile_seq.size());
  const size_t o_mtq = put(mem_tile_q0.data(), sizeof(int) * mem_tile_q0.size());
  const size_t o_dti = put(dec_tile_item.data(), sizeof(int) * dec_tile_item.size());
  const size_t o_dts = put(dec_tile_slot0.data(), sizeof(int) * dec_tile_slot0.size());
  p->dev_bytes = blob.size();
  size_t cap = 4096;
  while (cap < blob.size()) cap <<= 1;
  p->cap_bytes = cap;
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < h->plan_pool.size(); ++i)
    if (h->plan_pool[i].cap == cap) {
      p->dev = h->plan_pool[i].dev; p->host = h->plan_pool[i].host;
      h->plan_pool.erase(h->plan_pool.begin() + i);
      break;
    }
  if (!p->dev) {
    e = cudaMalloc(&p->dev, cap);
    if (e == cudaSuccess) e = cudaMallocHost(&p->host, cap);
  }
  if (e == cudaSuccess) {
    // pinned staging owned by the plan: the copy is truly asynchronous and needs no stream synchronisation.  A
    // recycled host block may still be the source of its previous plan's upload only if that upload has not run
    // yet -- it was enqueued at least 63 plans ago on a stream this handle has launched whole forwards on since.
    std::memcpy(p->host, blob.data(), blob.size());
    e = cudaMemcpyAsync(p->dev, p->host, blob.size(), cudaMemcpyHostToDevice, st);
  }
  if (e != cudaSuccess) {
    free_plan(p);
    return set_err(h, RESEP_ECUDA, std::string("plan upload: ") + cudaGetErrorString(e));
  }
  unsigned char* base = static_cast<unsigned char*>(p->dev);
  p->d_item_off = reinterpret_cast<const int64_t*>(base + o_off);
  p->d_item_len = reinterpret_cast<const int64_t*>(base + o_len);
  p->d_item_L = reinterpret_cast<const int*>(base + o_L);
  p->d_item_row0 = reinterpret_cast<const int*>(base + o_row0);
  p->d_chunk_item = reinterpret_cast<const int*>(base + o_ci);
  p->d_chunk_frame0 = reinterpret_cast<const int*>(base + o_cf);
  p->d_mem_pos = reinterpret_cast<const int*>(base + o_mp);
  p->d_mem_seq_off = reinterpret_cast<const int*>(base + o_ms);
  p->d_mem_tile_seq = reinterpret_cast<const int*>(base + o_mts);
  p->d_mem_tile_q0 = reinterpret_cast<const int*>(base + o_mtq);
  p->d_dec_tile_item = reinterpret_cast<const int*>(base + o_dti);
  p->d_dec_tile_slot0 = reinterpret_cast<const int*>(base + o_dts);
  p->last_use = h->tick;
  if (h->plans.size() >= 64) {   // evict the least recently used plan; only the graphs captured over ITS tables go with it
    auto it = std::min_element(h->plans.begin(), h->plans.end(),
                               [](const Plan* a, const Plan* b) { return a->last_use < b->last_use; });
    drop_graphs_of(h, *it);
    recycle_plan(h, *it);
    h->plans.erase(it);
  }
  h->plans.push_back(p);
  *out = p;
  return RESEP_OK;
}

// ------------------------------------------------------------------------------ workspace
struct Workspace {
  float *x0, *a, *o, *y, *qkv, *ctx, *hid, *hc_in, *hc_out;
  size_t bytes;
};

static Workspace carve(void* base, int64_t M, int64_t n_chunks) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(static_cast<unsigned char*>(base) + off) : nullptr;
    off = align_up(off + floats * sizeof(float), 1024);
    return p;
  };
  const size_t rows = (size_t)align_up((size_t)M, 128);   // GEMM tiles may touch a rounded-up row count
  w.x0 = take(rows * D);
  w.a = take(rows * D);
  w.o = take(rows * D);
  w.y = take(rows * D);
  w.qkv = take(rows * 3 * D);
  w.ctx = take(rows * D);
  w.hid = take(rows * FFN);
  w.hc_in = take(align_up((size_t)n_chunks, 128) * D);
  w.hc_out = take(align_up((size_t)n_chunks, 128) * D);
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------------------ forward
struct SeqDesc {          // the sequences one transformer block runs over
  int64_t rows;
  int n_seq;
  int seq_len;            // equal-length case (seq_off == nullptr)
  const int* seq_off;     // ragged case
  const int* pos;         // position of each row in its sequence (nullptr: row % seq_len)
  const int* tile_seq;    // attention query tiles of the ragged case
  const int* tile_q0;
  int n_tiles;
  int max_len;            // longest sequence
  bool intra = false;     // the sequences are the 150-row chunks of an intra block (not a memory sequence that happens to be 150 long)
};

static int run_layer(ResepHandle* h, const LayerDev& lw, float* o, const SeqDesc& sd, const Workspace& ws, int precision,
                     cudaStream_t st) {
  int rc;
  if (precision == RESEP_PREC_FP32) {
    if ((rc = launch_layernorm<float>(h, o, lw.norm1_w, lw.norm1_b, ws.y, sd.rows, st))) return rc;
    if ((rc = launch_gemm_f32(h, ws.y, lw.in_w, lw.in_b, nullptr, ws.qkv, sd.rows, 3 * D, D, false, st))) return rc;
    if ((rc = launch_attention_f32(h, ws.qkv, ws.ctx, sd.n_seq, sd.seq_len, sd.seq_off, sd.tile_seq, sd.tile_q0,
                                   sd.n_tiles, st))) return rc;
    if ((rc = launch_gemm_f32(h, ws.ctx, lw.out_w, lw.out_b, o, o, sd.rows, D, D, false, st))) return rc;
    if ((rc = launch_layernorm<float>(h, o, lw.norm2_w, lw.norm2_b, ws.y, sd.rows, st))) return rc;
    if ((rc = launch_gemm_f32(h, ws.y, lw.f1_w, lw.f1_b, nullptr, ws.hid, sd.rows, FFN, D, true, st))) return rc;
    if ((rc = launch_gemm_f32(h, ws.hid, lw.f2_w, lw.f2_b, o, o, sd.rows, D, FFN, false, st))) return rc;
    return RESEP_OK;
  }
  return tc_run_layer(h, lw, o, sd.rows, sd.n_seq, sd.seq_len, sd.seq_off, sd.tile_seq, sd.tile_q0, sd.n_tiles, sd.max_len, ws.y,
                      ws.qkv, ws.ctx, ws.hid, precision, st, sd.intra);
}

static int run_block(ResepHandle* h, int blk, const float* xprev, const float* hc, float* xin, float* o, float* out,
                     float* seq_mean, const SeqDesc& sd, const Workspace& ws, int precision, cudaStream_t st,
                     bf16* prelu_out = nullptr, bool prologue_done = false) {
  int rc;
  const BlockDev& bw = h->w.blk[blk];
  if (!prologue_done && (rc = launch_block_prologue(h, xprev, hc, xin, o, sd.rows, sd.pos, sd.seq_len, st))) return rc;
  for (int l = 0; l < NL; ++l)
    if ((rc = run_layer(h, bw.layers[l], o, sd, ws, precision, st))) return rc;
  return launch_block_epilogue(h, o, bw.fn_w, bw.fn_b, bw.gln_w, bw.gln_b, xin, out, seq_mean, sd.n_seq, sd.seq_len,
                               sd.seq_off, st, prelu_out, h->w.prelu_a);
}

static int forward_eager(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len, int B,
                         float* est, void* workspace, size_t workspace_bytes, int precision, int batch_mode,
                         cudaStream_t st, const ResepDebugOut* dbg, const ResepSpanCtl* span = nullptr);

static void drop_graphs(ResepHandle* h) { drop_graphs_of(h); }

static const Plan* find_plan(const ResepHandle* h, int B, const int64_t* item_off, const int64_t* item_len, int batch_mode) {
  for (const Plan* p : h->plans) {
    if (p->B != B || p->key.size() != 2 + 2 * (size_t)B || p->key[1] != batch_mode) continue;
    bool same = true;
    for (int i = 0; i < B && same; ++i) same = p->key[2 + 2 * i] == item_off[i] && p->key[3 + 2 * i] == item_len[i];
    if (same) return p;
  }
  return nullptr;
}

// Forward pass, replayed from a CUDA graph when the same (shapes, buffers, mode) has been seen before.
static int forward_impl(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len, int B,
                        float* est, void* workspace, size_t workspace_bytes, int precision, int batch_mode,
                        cudaStream_t st, const ResepDebugOut* dbg) {
  if (!h) return RESEP_EINVAL;
  if (dbg != nullptr || h->prof_on || !h->use_graphs || !mix || !item_off || !item_len || !est || B <= 0)
    return forward_eager(h, mix, item_off, item_len, B, est, workspace, workspace_bytes, precision, batch_mode, st, dbg);
  std::vector<uint64_t> key;
  key.reserve(6 + 2 * (size_t)B);
  key.push_back((uint64_t)(uintptr_t)mix); key.push_back((uint64_t)(uintptr_t)est); key.push_back((uint64_t)(uintptr_t)workspace);
  key.push_back((uint64_t)workspace_bytes); key.push_back((uint64_t)precision * 16 + (uint64_t)batch_mode); key.push_back((uint64_t)B);
  for (int i = 0; i < B; ++i) { key.push_back((uint64_t)item_off[i]); key.push_back((uint64_t)item_len[i]); }
  h->tick++;
  ResepHandle::GraphRec* rec = nullptr;
  for (auto& g : h->graphs) if (g.key == key) { rec = &g; break; }
  if (rec && rec->exec) {
    rec->last_use = h->tick;
    RESEP_CUDA(h, cudaSetDevice(h->device));
    RESEP_CUDA(h, cudaGraphLaunch(rec->exec, st));
    h->launches += rec->launches;
    return RESEP_OK;
  }
  if (!rec) {   // first sighting: run eagerly and remember the key
    int rc = forward_eager(h, mix, item_off, item_len, B, est, workspace, workspace_bytes, precision, batch_mode, st, nullptr);
    if (rc) return rc;
    if (h->graphs.size() >= 48) {
      auto it = std::min_element(h->graphs.begin(), h->graphs.end(), [](const ResepHandle::GraphRec& a, const ResepHandle::GraphRec& b) { return a.last_use < b.last_use; });
      if (it->exec) cudaGraphExecDestroy(it->exec);
      h->graphs.erase(it);
    }
    ResepHandle::GraphRec g;
    g.key = key; g.last_use = h->tick;
    g.plan = find_plan(h, B, item_off, item_len, batch_mode);
    h->graphs.push_back(g);
    return RESEP_OK;
  }
  // second sighting: capture on the handle's own stream, instantiate, launch on the caller's stream
  RESEP_CUDA(h, cudaSetDevice(h->device));
  if (!h->cap_stream) RESEP_CUDA(h, cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  const int64_t l0 = h->launches;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
  int rc = RESEP_OK;
  if (e == cudaSuccess) {
    rc = forward_eager(h, mix, item_off, item_len, B, est, workspace, workspace_bytes, precision, batch_mode, h->cap_stream, nullptr);
    e = cudaStreamEndCapture(h->cap_stream, &graph);
  }
  cudaGraphExec_t exec = nullptr;
  if (e == cudaSuccess && rc == RESEP_OK && graph) e = cudaGraphInstantiate(&exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (e != cudaSuccess || rc != RESEP_OK || !exec) {   // capture failed: run eagerly; give up on graphs after three failures
    cudaGetLastError();
    h->launches = l0;
    if (++h->graph_failures >= 3) h->use_graphs = 0;   // per handle: another handle's trouble does not switch this one's graphs off
    drop_graphs(h);
    return forward_eager(h, mix, item_off, item_len, B, est, workspace, workspace_bytes, precision, batch_mode, st, nullptr);
  }
  rec->exec = exec;
  rec->plan = find_plan(h, B, item_off, item_len, batch_mode);   // (the plan may have been re-created since the first sighting)
  rec->launches = h->launches - l0;
  rec->last_use = h->tick;
  RESEP_CUDA(h, cudaGraphLaunch(exec, st));
  return RESEP_OK;
}

static int forward_eager(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len, int B,
                         float* est, void* workspace, size_t workspace_bytes, int precision, int batch_mode,
                         cudaStream_t st, const ResepDebugOut* dbg, const ResepSpanCtl* span) {
  if (!h) return RESEP_EINVAL;
  if (!mix || !item_off || !item_len || !est || B <= 0) return set_err(h, RESEP_EINVAL, "null pointer or B <= 0");
  if (precision < RESEP_PREC_FP32 || precision > RESEP_PREC_FP16) return set_err(h, RESEP_EINVAL, "unknown precision");
  // RESEP_PREC_FP16 takes every code path of the bf16 mode: same kernels, instantiated for IEEE fp16 operands, with every
  // weight as hi + lo
  h->fmt16 = precision == RESEP_PREC_FP16;
  h->w16_mode = h->fmt16 ? h->w16_mode_fp16 : h->w16_mode_bf16;
  if (precision == RESEP_PREC_FP16) precision = RESEP_PREC_BF16;
  if (batch_mode != RESEP_BATCH_COUPLED && batch_mode != RESEP_BATCH_INDEPENDENT && !(span && batch_mode == RESEP_BATCH_SPAN_EXACT))
    return set_err(h, RESEP_EINVAL, "unknown batch_mode");
  // span != nullptr: one phase of a forward whose memory transformer runs elsewhere (resep_forward_span)
  const bool phase1 = span && span->phase == 1, phase2 = span && span->phase == 2;
  RESEP_CUDA(h, cudaSetDevice(h->device));
  Plan* p = nullptr;
  int rc = get_plan(h, B, item_off, item_len, batch_mode, st, &p);
  if (rc) return rc;
  Workspace ws = carve(workspace, p->M, p->n_chunks);
  if (!workspace || workspace_bytes < ws.bytes)
    return set_err(h, RESEP_EWORKSPACE, "workspace too small: need " + std::to_string(ws.bytes) + " bytes");
  if (precision != RESEP_PREC_FP32 && (rc = tc_init(h))) return rc;

  // The encoder also writes the first block's o = x0 + pe when that block runs as ONE slice (a sliced block re-uses the
  // same o scratch for every slice, so its prologue stays per slice).
  static const int64_t slice_env0 = getenv("RESEP_SLICE_CHUNKS") ? atoll(getenv("RESEP_SLICE_CHUNKS")) : 378;
  static const bool fuse_env = !(getenv("RESEP_ENC_FUSE") && getenv("RESEP_ENC_FUSE")[0] == '0');
  const bool enc_fused = fuse_env && !phase2 && (p->n_chunks <= 512 || slice_env0 <= 0);
  if (!phase2 && (rc = launch_encoder_chunked(h, mix, *p, ws.x0, st, enc_fused ? ws.o : nullptr))) return rc;
  if (dbg && dbg->enc) RESEP_CUDA(h, cudaMemcpyAsync(dbg->enc, ws.x0, p->M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));

  SeqDesc intra{p->M, (int)p->n_chunks, CHUNK, nullptr, nullptr, nullptr, nullptr, 0, CHUNK};
  SeqDesc mem{p->n_chunks, p->n_mem_seq, 0, p->d_mem_seq_off, p->d_mem_pos, p->d_mem_tile_seq, p->d_mem_tile_q0,
              p->n_mem_tiles, p->max_mem_len};
  if (p->n_mem_seq == 1) {   // a single sequence is the equal-length case
    mem.seq_len = (int)p->n_chunks; mem.seq_off = nullptr; mem.pos = nullptr; mem.tile_seq = nullptr; mem.tile_q0 = nullptr;
  }

  // The intra blocks are row-local per chunk, so large batches run them in SLICES of whole chunks: 378 chunks =
  // 56,700 rows = 221.5 pair-tiles fill the 74 CTA pairs for exactly three rounds (no wave-quantisation loss) and
  // keep a slice's residual stream + qkv + ctx (87 MB) inside the 126 MB L2; the per-slice scratch (o, qkv, ctx)
  // reuses the same addresses for every slice.  Batches of up to 512 chunks run as one slice.
  static const int64_t slice_env = getenv("RESEP_SLICE_CHUNKS") ? atoll(getenv("RESEP_SLICE_CHUNKS")) : 378;
  const int64_t slice = (p->n_chunks <= 512 || slice_env <= 0) ? p->n_chunks : slice_env;
  auto run_intra = [&](int blk, const float* xprev, const float* hc, float* xin, float* out, float* seq_mean,
                       bf16* prelu_out_, bool prologue_done = false) -> int {
    for (int64_t c0 = 0; c0 < p->n_chunks; c0 += slice) {
      const int64_t nc = std::min<int64_t>(slice, p->n_chunks - c0), r0 = c0 * CHUNK * D;
      SeqDesc sl{nc * CHUNK, (int)nc, CHUNK, nullptr, nullptr, nullptr, nullptr, 0, CHUNK, true};
      int rc2 = run_block(h, blk, xprev + r0, hc ? hc + c0 * D : nullptr, xin + r0, ws.o, out + r0,
                          seq_mean ? seq_mean + c0 * D : nullptr, sl, ws, precision, st, prelu_out_ ? prelu_out_ + r0 : nullptr, prologue_done);
      if (rc2) return rc2;
    }
    return RESEP_OK;
  };
  // seg_model[0](x + 0): skip input is the encoder output itself
  if (!phase2 && (rc = run_intra(0, ws.x0, nullptr, ws.x0, ws.a, ws.hc_in, nullptr, enc_fused))) return rc;
  if (phase1) {
    RESEP_CUDA(h, cudaMemcpyAsync(span->chunk_means, ws.hc_in, p->n_chunks * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return RESEP_OK;
  }
  if (dbg && dbg->seg0) RESEP_CUDA(h, cudaMemcpyAsync(dbg->seg0, ws.a, p->M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (dbg && dbg->chunk_mean)
    RESEP_CUDA(h, cudaMemcpyAsync(dbg->chunk_mean, ws.hc_in, p->n_chunks * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // mem_model[0](chunk means): memory transformer over chunk summaries
  if (phase2)
    RESEP_CUDA(h, cudaMemcpyAsync(ws.hc_out, span->hc, p->n_chunks * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else if ((rc = run_block(h, 2, ws.hc_in, nullptr, ws.hc_in, ws.o, ws.hc_out, nullptr, mem, ws, precision, st))) return rc;
  if (dbg && dbg->mem0)
    RESEP_CUDA(h, cudaMemcpyAsync(dbg->mem0, ws.hc_out, p->n_chunks * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // seg_model[1](out + hc)
  // in the bf16 mode the block epilogue also emits PReLU(out) in bf16 (ws.y), the A operand of output_fc
  // (the upper half of ws.y: its lower half stays per-slice scratch of the unfused layer path)
  bf16* prelu_out = precision == RESEP_PREC_BF16 ? reinterpret_cast<bf16*>(ws.y) + align_up((size_t)p->M, 128) * D : nullptr;
  if ((rc = run_intra(1, ws.a, ws.hc_out, ws.a, ws.a, nullptr, prelu_out))) return rc;
  if (dbg && dbg->seg1) RESEP_CUDA(h, cudaMemcpyAsync(dbg->seg1, ws.a, p->M * D * sizeof(float), cudaMemcpyDeviceToDevice, st));

  // output_fc (PReLU -> 1x1 conv 128->256) -> ReLU mask -> x encoder features -> decoder
  float* mask = ws.hid;
  if (precision == RESEP_PREC_FP32) {
    if ((rc = launch_prelu(h, ws.a, h->w.prelu_a, ws.y, p->M * D, st))) return rc;
    if ((rc = launch_gemm_f32(h, ws.y, h->w.fc_w, h->w.fc_b, nullptr, mask, p->M, NSPK * D, D, true, st))) return rc;
  } else {
    static const bool use_maskdec = !(getenv("RESEP_MASKDEC") && getenv("RESEP_MASKDEC")[0] == '0');
    if (prelu_out && p->maskdec_ok && use_maskdec) return launch_maskdec(h, prelu_out, ws.x0, *p, est, st);
    float* y_mask = prelu_out ? reinterpret_cast<float*>(prelu_out) : ws.y;
    if ((rc = tc_run_mask(h, ws.a, y_mask, mask, p->M, precision, st, prelu_out != nullptr))) return rc;
  }
  return launch_decoder(h, mask, ws.x0, *p, est, st);
}

}  // namespace resep

using namespace resep;

extern "C" {

int resep_create(const ResepConfig* cfg, const ResepWeights* w, int device, ResepHandle** out) {
  if (!cfg || !w || !out) return set_err(nullptr, RESEP_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->abi_version != RESEP_ABI_VERSION) return set_err(nullptr, RESEP_EINVAL, "ABI version mismatch");
  if (cfg->n_filters != D || cfg->kernel_size != KSZ || cfg->stride != STRIDE || cfg->segment_size != CHUNK ||
      cfg->n_heads != NH || cfg->d_ffn != FFN || cfg->n_layers != NL || cfg->n_blocks != 2 || cfg->n_spks != NSPK)
    return set_err(nullptr, RESEP_EINVAL,
                   "unsupported architecture: kernels are specialised for resepformer-wsj02mix "
                   "(128 filters, k16 s8, chunk 150, 8 heads, ffn 1024, 8 layers, 2 blocks, 2 speakers)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_err(nullptr, RESEP_ENODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return set_err(nullptr, RESEP_EINVAL, "device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return set_err(nullptr, RESEP_ECUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return set_err(nullptr, RESEP_ENODEVICE, "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                                 "; this library holds sm_100a code only");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return set_err(nullptr, RESEP_ECUDA, cudaGetErrorString(e));
  ResepHandle* h = new (std::nothrow) ResepHandle();
  if (!h) return set_err(nullptr, RESEP_EINVAL, "out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  if (const char* r = getenv("RESEP_RESERVE_SMS")) {   // persistent kernels leave this many SMs to a concurrent forward's small kernels
    const int k = atoi(r);
    if (k > 0 && k < h->sm_count - 2) h->sm_count -= k;
  }
  if (const char* m = getenv("RESEP_W16")) {   // weight operand of the bf16 mode (see DESIGN.md "precision modes")
    if (!strcmp(m, "bf16")) h->w16_mode = 0;          // single rounded bf16 weight: fails the SI-SNR gate
    else if (!strcmp(m, "bf16x2")) h->w16_mode = 1;   // every weight as hi + lo
    else if (!strcmp(m, "mixed")) h->w16_mode = 2;    // default
  }
  h->w16_mode_bf16 = h->w16_mode;
  if (const char* m = getenv("RESEP_W16F"))   // fp16 mode: which weights are hi + lo ("ffn2" / "ffn1": the attention projections and that FFN matrix)
    h->w16_mode_fp16 = !strcmp(m, "mixed") ? 2 : !strcmp(m, "single") ? 0 : !strcmp(m, "ffn2") ? 3 : !strcmp(m, "ffn1") ? 4 : 1;
  if (const char* g = getenv("RESEP_GRAPH")) h->use_graphs = g[0] != '0';
  int rc = upload_weights(h, w);
  if (rc) {
    g_create_err = h->err;
    if (h->arena) cudaFree(h->arena);
    delete h;
    return rc;
  }
  *out = h;
  return RESEP_OK;
}

int resep_load_weights(ResepHandle* h, const ResepWeights* w) {
  if (!h) return RESEP_EINVAL;
  RESEP_CUDA(h, cudaSetDevice(h->device));
  RESEP_CUDA(h, cudaDeviceSynchronize());
  drop_graphs_of(h);   // graphs bake in the kernel parameters (biases are passed by value) of the old weights
  int rc = upload_weights(h, w);
  if (rc) return rc;
  tc_destroy(h);   // packed tensor-core copies are rebuilt lazily from the new weights
  return RESEP_OK;
}

int resep_destroy(ResepHandle* h) {
  if (!h) return RESEP_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  tc_destroy(h);
  resep_profile(h, 0);
  drop_graphs_of(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (Plan* p : h->plans) free_plan(p);
  for (auto& b : h->plan_pool) { cudaFree(b.dev); cudaFreeHost(b.host); }
  if (h->arena) cudaFree(h->arena);
  delete h;
  return RESEP_OK;
}

const char* resep_last_error(const ResepHandle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int64_t resep_launch_count(const ResepHandle* h) { return h ? h->launches : 0; }

// development aid (not in the public header): copy the k_post2_tc clock trace of CTA 0 to `out[1536]`
extern "C" int resep_debug_trace(long long* out) {
  if (!resep::g_post_trace) return -1;
  cudaDeviceSynchronize();
  return cudaMemcpy(out, resep::g_post_trace, 1536 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
}

// development aid (not in the public header): the k_attn_tc clock trace of CTA 0, [2 roles][512] (tag, clock) pairs
extern "C" int resep_debug_attn_trace(long long* out) {
  if (!resep::g_attn_trace) return -1;
  cudaDeviceSynchronize();
  return cudaMemcpy(out, resep::g_attn_trace, 2048 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -3;
}

int resep_profile(ResepHandle* h, int enable) {
  if (!h) return RESEP_EINVAL;
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  h->prof.clear();
  h->prof_on = enable != 0;
  return RESEP_OK;
}

int resep_profile_report(ResepHandle* h, char* buf, size_t cap) {
  if (!h) return RESEP_EINVAL;
  if (!buf || cap < 3) return set_err(h, RESEP_EINVAL, "buffer too small");
  RESEP_CUDA(h, cudaSetDevice(h->device));
  RESEP_CUDA(h, cudaDeviceSynchronize());
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<long long> cnt;
  for (auto& r : h->prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) t = 0.f;
    size_t i = 0;
    for (; i < names.size(); ++i) if (names[i] == r.name) break;
    if (i == names.size()) { names.push_back(r.name); ms.push_back(0.0); cnt.push_back(0); }
    ms[i] += t;
    cnt[i] += 1;
  }
  std::string js = "{";
  for (size_t i = 0; i < names.size(); ++i) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", i ? ", " : "", names[i].c_str(), ms[i], cnt[i]);
    js += tmp;
  }
  js += "}";
  snprintf(buf, cap, "%s", js.c_str());
  return resep_profile(h, 0);
}

int resep_workspace_bytes(ResepHandle* h, int B, const int64_t* item_len, int precision, size_t* bytes) {
  if (!h) return RESEP_EINVAL;
  if (!item_len || !bytes || B <= 0) return set_err(h, RESEP_EINVAL, "null pointer or B <= 0");
  (void)precision;
  int64_t chunks = 0;
  for (int i = 0; i < B; ++i) {
    if (item_len[i] < KSZ)
      return set_err(h, RESEP_ESHORT, "item " + std::to_string(i) + " has " + std::to_string(item_len[i]) +
                                          " samples; kernel size (16) can't be greater than actual input size");
    chunks += ((item_len[i] - KSZ) / STRIDE + 1) / CHUNK + 1;
  }
  *bytes = carve(nullptr, chunks * CHUNK, chunks).bytes;
  return RESEP_OK;
}

int resep_forward(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len, int B, float* est,
                  void* workspace, size_t workspace_bytes, int precision, int batch_mode, void* stream) {
  return forward_impl(h, mix, item_off, item_len, B, est, workspace, workspace_bytes, precision, batch_mode,
                      static_cast<cudaStream_t>(stream), nullptr);
}

int resep_forward_span(ResepHandle* h, const float* mix, int64_t span_len, float* est, void* workspace, size_t workspace_bytes,
                       int precision, void* stream, const ResepSpanCtl* ctl) {
  if (!h) return RESEP_EINVAL;
  if (!ctl || (ctl->phase != 1 && ctl->phase != 2) || (ctl->phase == 1 && !ctl->chunk_means) || (ctl->phase == 2 && !ctl->hc))
    return set_err(h, RESEP_EINVAL, "forward_span: bad control block");
  const int64_t off = 0;
  return forward_eager(h, mix, &off, &span_len, 1, est, workspace, workspace_bytes, precision,
                       ctl->inner ? RESEP_BATCH_SPAN_EXACT : RESEP_BATCH_INDEPENDENT, static_cast<cudaStream_t>(stream), nullptr, ctl);
}

int resep_memory_workspace_bytes(ResepHandle* h, int n_chunks, size_t* bytes) {
  if (!h) return RESEP_EINVAL;
  if (!bytes || n_chunks <= 0) return set_err(h, RESEP_EINVAL, "null pointer or n_chunks <= 0");
  *bytes = carve(nullptr, n_chunks, n_chunks).bytes;
  return RESEP_OK;
}

int resep_memory_block(ResepHandle* h, const float* chunk_means, float* hc, int n_chunks, void* workspace, size_t workspace_bytes,
                       int precision, void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!chunk_means || !hc || n_chunks <= 0) return set_err(h, RESEP_EINVAL, "null pointer or n_chunks <= 0");
  if (precision < RESEP_PREC_FP32 || precision > RESEP_PREC_FP16) return set_err(h, RESEP_EINVAL, "unknown precision");
  h->fmt16 = precision == RESEP_PREC_FP16;
  h->w16_mode = h->fmt16 ? h->w16_mode_fp16 : h->w16_mode_bf16;
  if (precision == RESEP_PREC_FP16) precision = RESEP_PREC_BF16;
  if (n_chunks > h->w.pe_rows) return set_err(h, RESEP_EPOS, "memory sequence longer than the positional-encoding table");
  RESEP_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws = carve(workspace, n_chunks, n_chunks);
  if (!workspace || workspace_bytes < ws.bytes)
    return set_err(h, RESEP_EWORKSPACE, "workspace too small: need " + std::to_string(ws.bytes) + " bytes");
  int rc;
  if (precision != RESEP_PREC_FP32 && (rc = tc_init(h))) return rc;
  RESEP_CUDA(h, cudaMemcpyAsync(ws.hc_in, chunk_means, (size_t)n_chunks * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SeqDesc mem{n_chunks, 1, n_chunks, nullptr, nullptr, nullptr, nullptr, 0, n_chunks};
  if ((rc = run_block(h, 2, ws.hc_in, nullptr, ws.hc_in, ws.o, ws.hc_out, nullptr, mem, ws, precision, st))) return rc;
  RESEP_CUDA(h, cudaMemcpyAsync(hc, ws.hc_out, (size_t)n_chunks * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return RESEP_OK;
}

int resep_resample_fir(ResepHandle* h, const float* x, int rows, int64_t n_in, float* y, int64_t n_out, int channels, int down,
                       int up, const float* taps, int ktaps, int width, void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!x || !y || !taps || rows < 0 || n_in < 0 || n_out < 0 || channels < 1 || down < 1 || up < 1 || ktaps < 1 || width < 0)
    return set_err(h, RESEP_EINVAL, "resample: bad argument");
  RESEP_CUDA(h, cudaSetDevice(h->device));
  return launch_resample_fir(h, x, rows, n_in, y, n_out, channels, down, up, taps, ktaps, width, static_cast<cudaStream_t>(stream));
}

int resep_peak_normalize(ResepHandle* h, float* est, const int64_t* item_off, const int64_t* item_len, int B, float* peaks,
                         void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!est || !item_off || !item_len || !peaks || B <= 0) return set_err(h, RESEP_EINVAL, "null pointer or empty batch");
  RESEP_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan* p = nullptr;
  for (Plan* q : h->plans) {                          // the forward that produced est left a plan with these offsets
    if (q->B != B || q->key.size() != 2 + 2 * (size_t)B) continue;
    bool same = true;
    for (int i = 0; i < B && same; ++i) same = q->key[2 + 2 * i] == item_off[i] && q->key[3 + 2 * i] == item_len[i];
    if (same) { p = q; break; }
  }
  int rc;
  if (!p && (rc = get_plan(h, B, item_off, item_len, RESEP_BATCH_INDEPENDENT, st, &p))) return rc;
  int64_t max_len = 0;
  for (int i = 0; i < B; ++i) max_len = item_len[i] > max_len ? item_len[i] : max_len;
  return launch_peak_normalize(h, est, *p, max_len, peaks, st);
}

int resep_forward_debug(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len, int B,
                        float* est, void* workspace, size_t workspace_bytes, int precision, int batch_mode, void* stream,
                        const ResepDebugOut* dbg) {
  return forward_impl(h, mix, item_off, item_len, B, est, workspace, workspace_bytes, precision, batch_mode,
                      static_cast<cudaStream_t>(stream), dbg);
}

int resep_encoder_fwd(ResepHandle* h, const float* mix, int64_t T, float* tokens_out, void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!mix || !tokens_out) return set_err(h, RESEP_EINVAL, "null pointer");
  if (T < KSZ) return set_err(h, RESEP_ESHORT, "kernel size (16) can't be greater than actual input size");
  RESEP_CUDA(h, cudaSetDevice(h->device));
  return launch_encoder_single(h, mix, T, tokens_out, static_cast<cudaStream_t>(stream));
}

int resep_layer_fwd(ResepHandle* h, int block, int layer, float* x, int n_seq, int seq_len, void* workspace,
                    size_t workspace_bytes, int precision, void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!x || block < 0 || block > 2 || layer < 0 || layer >= NL || n_seq <= 0 || seq_len <= 0)
    return set_err(h, RESEP_EINVAL, "bad argument");
  if (precision < RESEP_PREC_FP32 || precision > RESEP_PREC_FP16) return set_err(h, RESEP_EINVAL, "unknown precision");
  h->fmt16 = precision == RESEP_PREC_FP16;
  h->w16_mode = h->fmt16 ? h->w16_mode_fp16 : h->w16_mode_bf16;
  if (precision == RESEP_PREC_FP16) precision = RESEP_PREC_BF16;
  RESEP_CUDA(h, cudaSetDevice(h->device));
  const int64_t rows = (int64_t)n_seq * seq_len;
  Workspace ws = carve(workspace, rows, 0);
  if (!workspace || workspace_bytes < ws.bytes)
    return set_err(h, RESEP_EWORKSPACE, "workspace too small: need " + std::to_string(ws.bytes) + " bytes");
  int rc;
  if (precision != RESEP_PREC_FP32 && (rc = tc_init(h))) return rc;
  const int blk = block == 2 ? 2 : block;
  SeqDesc sd{rows, n_seq, seq_len, nullptr, nullptr, nullptr, nullptr, 0, seq_len, block < 2 && seq_len == CHUNK};
  return run_layer(h, h->w.blk[blk].layers[layer], x, sd, ws, precision, static_cast<cudaStream_t>(stream));
}

int resep_layer_kernel_repeat(ResepHandle* h, int block, int layer, int which, float* x, int n_seq, int seq_len,
                              void* workspace, size_t workspace_bytes, int precision, int reps, int pdl, void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!x || block < 0 || block > 2 || layer < 0 || layer >= NL || n_seq <= 0 || seq_len <= 0 || which < 0 || which > 2 || reps <= 0)
    return set_err(h, RESEP_EINVAL, "bad argument");
  if (precision != RESEP_PREC_BF16 && precision != RESEP_PREC_FP16)
    return set_err(h, RESEP_EINVAL, "resep_layer_kernel_repeat: the fused layer kernels exist in the bf16 and fp16 modes");
  h->fmt16 = precision == RESEP_PREC_FP16;
  h->w16_mode = h->fmt16 ? h->w16_mode_fp16 : h->w16_mode_bf16;
  RESEP_CUDA(h, cudaSetDevice(h->device));
  const int64_t rows = (int64_t)n_seq * seq_len;
  Workspace ws = carve(workspace, rows, 0);
  if (!workspace || workspace_bytes < ws.bytes)
    return set_err(h, RESEP_EWORKSPACE, "workspace too small: need " + std::to_string(ws.bytes) + " bytes");
  int rc;
  if ((rc = tc_init(h))) return rc;
  return tc_repeat_layer_kernel(h, h->w.blk[block].layers[layer], which, x, rows, n_seq, seq_len, ws.qkv, ws.ctx, reps, pdl != 0,
                                static_cast<cudaStream_t>(stream), block < 2 && seq_len == CHUNK);
}

int resep_linear_fwd(ResepHandle* h, const float* A, const float* W, const float* bias, float* out, int64_t M, int N,
                     int K, int relu, int precision, void* stream) {
  if (!h) return RESEP_EINVAL;
  if (!A || !W || !bias || !out || M <= 0) return set_err(h, RESEP_EINVAL, "bad argument");
  RESEP_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == RESEP_PREC_FP32) return launch_gemm_f32(h, A, W, bias, nullptr, out, M, N, K, relu != 0, st);
  int rc = tc_init(h);
  if (rc) return rc;
  return tc_linear_test(h, A, W, bias, out, M, N, K, relu != 0, precision, st);
}

}  // extern "C"
