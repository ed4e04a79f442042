// One persistent kernel per Llama decode step for 2..16 rows (inputs x beams): the beam-search path, which is what
// every shipped caller of the reference runs (procyon/evaluate/framework/procyon.py:71-76, beam_size = 2 x captions).
//
// Same skeleton as decode_megakernel.cu (one CTA per SM, the CTA's share of all 4 L + 1 weight matrices streamed through
// a shared-memory ring that runs ahead of the grid barriers, phases separated by grid-wide barriers), but built for
// up to 16 activation rows and without a producer warp (8 warps = 255 registers per thread; each warp refills the
// ring slots it owns right after it has drained them):
//   * a ring slot is a TILE of 16 weight rows x 256 k (4 TMA boxes of 16 x 64, 128-byte swizzle), so that one
//     ldmatrix.x4 of weights (A) and one of activations (B) feed two real mma.sync.m16n8k16 (16 weight rows x 16 rows);
//   * work is cut stream-K style: the chunk sequence (k-part, row group, k-chunk) of a matrix is split evenly over the
//     CTAs and, inside a CTA, over the 8 warps, whatever the row count of the matrix (the two short matrices, qkv and
//     o_proj, are cut at row-group boundaries instead: a split group costs two L2 round trips at phase end).  A warp keeps the
//     16 x 16 accumulator of its current row group in registers; pieces of a row group that end up in different warps
//     meet in a shared-memory pool, pieces in different CTAs in a global scratch with a ticket per output group - the
//     last contributor adds them in a fixed order (bit-reproducible) and runs the epilogue;
//   * the activations of a phase are staged once per CTA (rows x K bf16, padded rows: conflict-free ldmatrix); K of the
//     down projection does not fit, so that matrix is also cut into k-parts of <= d_model columns (the ticket merge
//     adds the parts);
//   * attention: all beams of an input share the prompt K/V and, mostly, their ancestors' generated K/V, so one work
//     item = (input, kv head, 64 keys) serves every beam: keys are the prompt positions followed by ALL generated
//     (step, physical row) entries of the input, with a per-key bit mask of the beams whose ancestry holds that entry;
//     S = Q K^T and O = P V run on mma.sync for beams x 4 query rows at once.  Partials are merged by (row, head)
//     units spread over all warps of the grid after a grid barrier.
//
// Replaces the per-token HF LlamaForCausalLM forward of _generate_beam_search (procyon/model/model_unified.py:769 ->
// procyon/model/pmc_llama.py:581) plus the per-layer KV reorder (model_unified.py:830-832) for beam_size x inputs rows.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace pcy {

namespace {

constexpr int RW = 8;             // consumer warps
constexpr int RT = RW * 32;       // consumer threads
constexpr int RBLOCK = RT;        // no producer warp: every warp refills its own ring slots (255 registers per thread)
constexpr int TR = 16;            // weight rows per tile
constexpr int TK = 256;           // k per tile
constexpr int SLOT_BYTES = TR * TK * 2;
constexpr int HD = 128;
constexpr int GQ = 4;
constexpr int PW = 3;             // pool tiles per consumer warp
constexpr int SUBK = 64;          // keys per attention work item
constexpr int KVP = 272;          // bytes per K / V / Q row in shared memory (256 + 16: conflict-free ldmatrix)
constexpr int PPB = 144;          // bytes per P row (64 bf16 + 16)
constexpr int PSTR = HD + 4;      // floats per (split, head) attention partial: 128 outputs, max, sum
constexpr int MERGE_B = 18;       // splits merged per batch of independent loads (72 registers)
constexpr int KV_ISSUERS = 2 * SUBK;  // threads that fetch K / V rows of an item (arrival count of its mbarrier)
// misc block of shared memory: [0, 96) row group of every pool tile | [128, 192) tokens | [192, 232) chunk ranges |
// [256, ...) mbarriers (full /
// empty per ring slot + one for the attention tiles) | [MISC_LN, ...) two RMSNorm weight pointers per layer
constexpr int MAX_SLOTS = 40;
constexpr int MISC_TOK = 128, MISC_LOHI = 192, MISC_BARS = 256, MISC_LN = MISC_BARS + 8 * (MAX_SLOTS + 1) + 8;

enum : int { EPI_BF16 = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_FP32 = 3 };
enum : int { STAGE_PLAIN = 0, STAGE_RMS = 1 };

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int atom_release_add(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory"); }
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const void* desc, uint32_t bar, int32_t c0, int32_t c1,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
  return u;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 u) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Grid-wide barrier over the consumer threads of all CTAs (cooperative launch); see decode_megakernel.cu.
struct GridBarrier {
  unsigned int* counter;
  unsigned int target;
  unsigned int nblocks;
  __device__ __forceinline__ void sync() {
    consumer_sync();
    if (threadIdx.x == 0) {
      target += nblocks;
      red_release_add(counter, 1u);
      uint64_t t0 = 0;
      for (uint32_t it = 0; ld_relaxed_u32(counter) < target; ++it) {
        if ((it & 0x3fffu) == 0x3fffu) {  // bounded: a scheduling bug must trap, not hang the GPU
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ull) __trap();
        }
      }
    }
    consumer_sync();
  }
};

// ---- how a weight matrix [N, K] is cut ----------------------------------------------------------------------------
// k-chunks of 256; K is cut into KQ parts of ckq chunks (the last may be shorter) so that one part of the activations
// fits the staging area; row groups of 16.  Chunk sequence: for part q: for row group rg: for k-chunk in part q.
// CTA b takes chunks [T b / n, T (b + 1) / n).
struct Geom {
  int n_rg, ck, KQ, ckq, per, T;
  int unit;  // CTA ranges begin at multiples of `unit` chunks: 1 = anywhere (stream-K), ck = at row-group boundaries
};
__host__ __device__ inline Geom make_geom(int N, int K, int kcap, bool aligned = false) {
  Geom g;
  g.n_rg = (N + TR - 1) / TR;
  g.ck = K / TK;
  const int kq0 = (K + kcap - 1) / kcap;
  g.ckq = (g.ck + kq0 - 1) / kq0;
  g.KQ = (g.ck + g.ckq - 1) / g.ckq;
  g.per = g.n_rg * g.ckq;
  g.T = g.n_rg * g.ck;
  g.unit = (aligned && g.KQ == 1) ? g.ck : 1;
  return g;
}
// Matrices whose phases are short (qkv, o_proj) are cut at row-group boundaries instead: the largest share grows by
// up to one group (16 chunks of ~0.2 us), but no group is split between CTAs, and a split group costs its finisher
// two L2 round trips (ticket, pieces) at the very end of the phase, when every other CTA waits at the grid barrier.
__host__ __device__ inline int cta_lo(const Geom& g, int i, int nb) {
  return (int)((unsigned int)(g.T / g.unit) * (unsigned int)i / (unsigned int)nb) * g.unit;
}
__host__ __device__ inline int part_len(const Geom& g, int q) { return q == g.KQ - 1 ? g.ck - q * g.ckq : g.ckq; }
__host__ __device__ inline int seg_start(const Geom& g, int q, int rg) { return q * g.per + rg * part_len(g, q); }
// CTA that owns chunk c when T chunks are cut into nb ranges [T i / nb, T (i + 1) / nb)
// (32-bit arithmetic: the host plan requires T (nb + 1) < 2^31)
__host__ __device__ inline int chunk_owner(int c, int T, int nb) {
  return (int)(((unsigned int)(c + 1) * (unsigned int)nb - 1u) / (unsigned int)T);
}
__host__ __device__ inline int range_lo(int T, int i, int nb) {
  return (int)((unsigned int)T * (unsigned int)i / (unsigned int)nb);
}

struct RowsParams {
  pcy_llama_config cfg;
  const bf16* embed;
  const bf16* norm;
  const LlamaLayerPtrs* layers;
  const CUtensorMap* maps;  // [4 L + 1] device array: wqkv, wo, wgu, wdown of every layer, then the LM head
  const float* rope;
  int rows, beams, S, max_gen;
  const bf16* kv_prompt;
  const uint8_t* prompt_valid;
  bf16* kv_gen;
  const int32_t* tokens;
  const int32_t* slots;
  const int32_t* state;
  float* logits;
  bf16 *x, *qkv, *attn, *act;
  float* part;            // attention partials [rows][KVH][max_splits][GQ][PSTR]
  int max_splits;
  float* pieces;          // cross-CTA partial tiles [(og * nseg + seg) * maxcp + piece][16 x 8 NT]
  unsigned int* tickets;  // [max output groups], zero between phases
  unsigned int* barrier;
  int ring_slots;
  int kcap;      // activation columns the staging area holds per row (multiple of 256)
  int maxcp;     // piece slots per segment in `pieces`
  int off_pool;  // byte offsets inside the work area (after ring + barrier block)
  int off_misc;
  unsigned long long* timing;
  int timing_cta;
  Geom geom[5];    // per phase kind (0 qkv, 1 o_proj, 2 gate/up, 3 down, 4 LM head), computed on the host: the kernel
                   // is ~200 KB of SASS that runs out of a cold instruction cache at every phase boundary, and each
                   // inlined copy of the geometry arithmetic (five integer divisions) is 150 instructions of it
  int align_mask;  // bit k: phases of kind k (0 qkv, 1 o_proj, 2 gate/up, 4 LM head) are cut at row-group boundaries
};

struct PhaseDesc {
  int N, K, stage, epi, map, kind;
  const bf16* A;
  int64_t lda;
  const bf16* rms_w;
  void* out;
  int64_t ldo;
};

// phases: 4 l + {0 qkv, 1 o_proj, 2 gate/up, 3 down}, then the LM head
__device__ __forceinline__ PhaseDesc phase_desc(const RowsParams& p, const bf16* const* s_ln, int ph) {
  const pcy_llama_config& c = p.cfg;
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
  const int qkv_dim = (H + 2 * KVH) * HD;
  const int n_phases = 4 * c.n_layers + 1;
  PhaseDesc z;
  z.rms_w = nullptr;
  z.map = ph;
  z.kind = ph == n_phases - 1 ? 4 : (ph & 3);
  if (ph == n_phases - 1) {
    z.N = c.vocab; z.K = d; z.stage = STAGE_RMS; z.epi = EPI_FP32; z.A = p.x; z.lda = d; z.rms_w = p.norm;
    z.out = p.logits; z.ldo = c.vocab;
    return z;
  }
  const bf16* const* ln = s_ln + 2 * (ph >> 2);
  switch (ph & 3) {
    case 0:
      z.N = qkv_dim; z.K = d; z.stage = STAGE_RMS; z.epi = EPI_BF16; z.A = p.x; z.lda = d; z.rms_w = ln[0];
      z.out = p.qkv; z.ldo = qkv_dim;
      break;
    case 1:
      z.N = d; z.K = H * HD; z.stage = STAGE_PLAIN; z.epi = EPI_RESIDUAL; z.A = p.attn; z.lda = H * HD; z.out = p.x;
      z.ldo = d;
      break;
    case 2:
      z.N = 2 * f; z.K = d; z.stage = STAGE_RMS; z.epi = EPI_SWIGLU; z.A = p.x; z.lda = d; z.rms_w = ln[1];
      z.out = p.act; z.ldo = f;
      break;
    default:
      z.N = d; z.K = f; z.stage = STAGE_PLAIN; z.epi = EPI_RESIDUAL; z.A = p.act; z.lda = f; z.out = p.x; z.ldo = d;
      break;
  }
  return z;
}

// ---- the weight tiles of one warp, in consumption order ----------------------------------------------------------------
// There is no producer warp (a ninth warp would cap every thread at 168 registers): each warp owns ring_slots / 8 slots
// of the ring and refills a slot itself, right after it has consumed it, with the tile it will need ring-depth tiles
// later - whatever phase that tile belongs to, so weights keep streaming across grid barriers, staging and attention.
// The cursor walks the warp's tiles: phases -> runs (the CTA's chunks of one k-part) -> the warp's span of the run.
struct Cursor {
  int ph;              // phase of the next tile (n_phases = exhausted)
  int hi, b;           // end of the CTA's range in the phase, end of the current run
  int c, c_hi;         // next chunk / end of this warp's span in the run
  int q, len, qbase;   // run constants: k-part, chunks per (row group, part), first chunk of the part
  int ckq;
  int n_rg, ck, KQ, per, T;
  const int* lohi;     // shared memory: [kind][2] first / end chunk of this CTA's range for every phase kind

  __device__ __forceinline__ void begin_run(int a, int warp) {
    q = min(a / per, KQ - 1);
    b = min(hi, q == KQ - 1 ? T : (q + 1) * per);
    len = q == KQ - 1 ? ck - q * ckq : ckq;
    qbase = q * per;
    const int C = b - a;
    c = a + C * warp / RW;
    c_hi = a + C * (warp + 1) / RW;
  }
  __device__ __forceinline__ void begin_phase(const RowsParams& p, int warp) {
    const int kind = ph == 4 * p.cfg.n_layers ? 4 : (ph & 3);
    const Geom& g = p.geom[kind];
    n_rg = g.n_rg; ck = g.ck; KQ = g.KQ; ckq = g.ckq; per = g.per; T = g.T;
    hi = lohi[2 * kind + 1];
    b = lohi[2 * kind];
    c = c_hi = 0;
  }
  // the next tile: tensor map index, first weight row, first k element; false when every phase is exhausted
  __device__ __forceinline__ bool next(const RowsParams& p, int warp, int n_phases, int& map, int& row0, int& k0) {
    while (true) {
      if (ph >= n_phases) return false;
      if (c < c_hi) {
        const int rem = c - qbase;
        const int rgi = rem / len;
        map = ph;
        row0 = rgi * TR;
        k0 = (q * ckq + (rem - rgi * len)) * TK;
        ++c;
        return true;
      }
      if (b < hi) {
        begin_run(b, warp);
      } else {
        ++ph;
        if (ph < n_phases) begin_phase(p, warp);
      }
    }
  }
};

// ---- consumer side ---------------------------------------------------------------------------------------------------
struct Ctx {
  uint32_t ring, bars;  // this WARP's slots: slot i at ring + i * SLOT_BYTES, its mbarrier at bars + 8 i
  int depth;            // slots per warp
  uint32_t n_used;      // tiles consumed so far by this warp (tile e lives in slot e % depth, parity (e / depth) & 1)
  Cursor fill;          // the next tile to request
  int n_phases;
  uint32_t act;        // staged activations: row m at act + m * pitch
  uint32_t pitch;      // kcap * 2 + 16 bytes
  float* pool;         // [RW * PW][16][rs] partial tiles
  int rs;              // activation rows rounded up to an even number: columns of a partial tile
  int* prg;            // [RW * PW] row group of every pool tile (-1 = free)
  unsigned long long* tbuf;
  int tix, tcta;
  __device__ __forceinline__ void stamp() {
    if (tbuf != nullptr) {
      if (blockIdx.x == tcta && threadIdx.x == 0) tbuf[tix] = globaltimer_ns();
      ++tix;
    }
  }
};

// Request the warp's next tile into ring slot `slot` (just drained by this warp, or still untouched).
__device__ __forceinline__ void refill(const RowsParams& p, Ctx& cx, uint32_t slot) {
  int map, row0, k0;
  const int warp = threadIdx.x >> 5;
  if (!cx.fill.next(p, warp, cx.n_phases, map, row0, k0)) return;
  if ((threadIdx.x & 31) == 0) {
    fence_proxy_async_smem();  // the warp's ldmatrix reads of the slot are ordered before the bulk write
    const uint32_t bar = cx.bars + 8u * slot;
    mbar_arrive_expect_tx(bar, SLOT_BYTES);
    const uint32_t dst = cx.ring + slot * SLOT_BYTES;
    const uint64_t pol = l2_evict_first_policy();  // weights are read once per step: keep L2 for KV / activations
#pragma unroll
    for (int i = 0; i < 4; ++i) tma_load_2d_hint(dst + i * 2048, p.maps + map, bar, k0 + i * 64, row0, pol);
  }
}

// one tile (16 weight rows x 256 k) against the staged activations: 16 k-steps of ldmatrix(A) + ldmatrix(B) + NT mma
template <int NT>
__device__ __forceinline__ void mma_tile(uint32_t wt, uint32_t a_lane, float (&acc)[NT][4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t r = (lane & 7) + ((lane >> 3) & 1) * 8;  // weight row of this lane's ldmatrix address
  const uint32_t hi = lane >> 4;                          // k half (8 elements)
  const uint32_t w_row = wt + r * 128;
  const uint32_t sw = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    uint32_t af[4];
    const uint32_t chunk = ((ks & 3) << 1) + hi;
    ldsm_x4(w_row + (ks >> 2) * 2048 + ((chunk ^ sw) << 4), af);
    if (NT == 2) {
      uint32_t bfr[4];
      ldsm_x4(a_lane + ks * 32, bfr);
      mma16816(acc[0], af, bfr[0], bfr[1]);
      mma16816(acc[NT - 1], af, bfr[2], bfr[3]);
    } else {
      uint32_t b0, b1;
      ldsm_x2(a_lane + ks * 32, b0, b1);
      mma16816(acc[0], af, b0, b1);
    }
  }
}

// one output element
__device__ __forceinline__ void store_out(void* out, int64_t ldo, int epi, int rows, int n_out, int col, int m, float v) {
  if (m >= rows || col >= n_out) return;
  if (epi == EPI_FP32) {
    reinterpret_cast<float*>(out)[(int64_t)m * ldo + col] = v;
  } else {
    bf16* op = reinterpret_cast<bf16*>(out) + (int64_t)m * ldo + col;
    if (epi == EPI_RESIDUAL) v += ldcg_bf16(op);
    *op = __float2bfloat16_rn(v);
  }
}

// Weight phase: out = epi(W . A) for this CTA's chunk range.
// `after_consume` runs right after the CTA has consumed its last tile of a run (the staging area is free from then on):
// the qkv phase uses it to request the K / V rows of its first attention item before it turns to its epilogue.
template <int NT, class AfterConsume>
__device__ __forceinline__ void weight_phase(const RowsParams& p, Ctx& cx, const PhaseDesc& z, const int32_t* tok_rows,
                                             AfterConsume&& after_consume) {
  // partial tile: [16 weight rows][rs activation rows] fp32; when tiles are added, lane l owns weight row l / 2 and
  // the activation rows (l & 1) * rs / 2 ... of it
  constexpr int EPL = 4 * NT;       // upper bound of rs / 2
  const int RS = cx.rs, TS = TR * RS, eh = RS >> 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = gridDim.x, bid = blockIdx.x;
  const int rows = p.rows;
  const Geom g = p.geom[z.kind];
  const bool swiglu = z.epi == EPI_SWIGLU;
  const int n_out = swiglu ? z.N / 2 : z.N;
  const int lo = cx.fill.lohi[2 * z.kind], hi = cx.fill.lohi[2 * z.kind + 1];
  const int nseg = swiglu ? 2 : g.KQ;
  const unsigned int og_total = (unsigned int)(swiglu ? 2 * g.ck : g.ck);

  for (int a = lo; a < hi;) {
    const int q = min(a / g.per, g.KQ - 1);
    const int b = min(hi, q == g.KQ - 1 ? g.T : (q + 1) * g.per);
    const int len = part_len(g, q), qbase = q * g.per;
    const int k0 = q * g.ckq * TK, kn = len * TK;  // staged columns

    // ---- stage the activations of this k-part: rows m < rows, columns [k0, k0 + kn) ----
    // Every L2 round trip costs 2-3 us while all SMs stream weights, so the loads of ALL rows are issued before the
    // first use: thread t owns the 16-byte column pieces t and t + 256 of every row (and of the RMSNorm weight).
    if (tid < RW * PW) cx.prg[tid] = -1;
    if (rows <= 4)
    {
      // Up to 4 rows: cp.async (L2 -> shared memory, no registers in between) and rolled loops.  This block runs once
      // per phase out of a cold instruction cache (the kernel is ~200 KB of SASS), and with few rows the unrolled
      // register-staged form below costs more in instruction fetch than in data: 3.9 us for 2 rows against 1.7 us for
      // a plain copy of the same bytes (2 rows: 3.51 -> 3.40 ms per step).  From 5 rows on the three passes over shared
      // memory (raw copy, sums of squares, in-place normalisation) cost more than they save (10 rows: 7.5 vs 5.8 us).
      const int pcs = kn >> 3;  // 16-byte pieces per row
      const bool rms = z.stage == STAGE_RMS;
      for (int m = 0; m < rows; ++m) {
        const bf16* src = (tok_rows ? p.embed + (int64_t)tok_rows[m] * z.K : z.A + (int64_t)m * z.lda) + k0;
        for (int pc = tid; pc < pcs; pc += RT)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cx.act + m * cx.pitch + pc * 16),
                       "l"(src + pc * 8) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;  // RMSNorm weight of this thread's two column pieces
      if (rms) {
        if (tid < pcs) g0 = *reinterpret_cast<const uint4*>(z.rms_w + tid * 8);
        if (tid + RT < pcs) g1 = *reinterpret_cast<const uint4*>(z.rms_w + (tid + RT) * 8);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (rms) {
        // RMSNorm with HF rounding, in place: y = w * bf16(x * rstd), rstd from fp32 sums over the whole row.  Every
        // thread only touches the pieces it copied itself (t and t + 256 of every row: K <= 4096, checked on the host).
        float* red = cx.pool;  // [16][RW] sums of squares (the pool is idle while staging)
        for (int m = 0; m < rows; ++m) {
          float ss = 0.f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pc = tid + h * RT;
            if (pc < pcs) {
              const uint4 u = lds_v4(cx.act + m * cx.pitch + pc * 16);
              const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
              ss += a0.x * a0.x + a0.y * a0.y + a1.x * a1.x + a1.y * a1.y + a2.x * a2.x + a2.y * a2.y + a3.x * a3.x +
                    a3.y * a3.y;
            }
          }
          ss = warp_sum(ss);
          if (lane == 0) red[m * RW + warp] = ss;
        }
        consumer_sync();
        for (int m = 0; m < rows; ++m) {
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < RW; ++w) tot += red[m * RW + w];
          const float rstd = rsqrtf(tot / (float)z.K + p.cfg.rms_eps);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pc = tid + h * RT;
            if (pc < pcs) {
              const uint32_t addr = cx.act + m * cx.pitch + pc * 16;
              const uint4 u = lds_v4(addr);
              const uint4 gw = h == 0 ? g0 : g1;
              const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
              const uint32_t gg[4] = {gw.x, gw.y, gw.z, gw.w};
              uint32_t oo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // two elements at a time: bf16(x * rstd) by one packed conversion, times the weight by one packed
                // bf16 multiply (the product of two bf16 is exact in fp32, so rounding it once is what the fp32 form
                // w * bf16_round(x * rstd) -> bf16 gives)
                const float2 xv = unpack_bf16x2(uu[e]);
                const uint32_t t2 = pack_bf16x2(xv.x * rstd, xv.y * rstd);
                asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(oo[e]) : "r"(gg[e]), "r"(t2));
              }
              sts_v4(addr, make_uint4(oo[0], oo[1], oo[2], oo[3]));
            }
          }
        }
      }
    }
    else
    {
      constexpr int RBATCH = 10;  // rows per batch of loads (2 pieces per thread and row: 80 registers)
      const int pcs = kn >> 3;    // 16-byte pieces per row
      const bool rms = z.stage == STAGE_RMS;
      float* red = cx.pool;       // [16][RW] sums of squares (the pool is idle while staging)
      for (int pb = 0; pb < pcs; pb += 2 * RT) {  // (one iteration for K <= 4096; RMSNorm needs K <= 4096: checked on the host)
        const int pc0 = pb + tid, pc1 = pb + RT + tid;
        const bool h0 = pc0 < pcs, h1 = pc1 < pcs;
        uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;
        if (rms) {
          if (h0) g0 = *reinterpret_cast<const uint4*>(z.rms_w + pc0 * 8);
          if (h1) g1 = *reinterpret_cast<const uint4*>(z.rms_w + pc1 * 8);
        }
        for (int mb = 0; mb < rows; mb += RBATCH) {
          uint4 u[RBATCH][2];
#pragma unroll
          for (int i = 0; i < RBATCH; ++i) {
            const int m = mb + i;
            u[i][0] = u[i][1] = make_uint4(0, 0, 0, 0);
            if (m < rows) {
              const bf16* src = (tok_rows ? p.embed + (int64_t)tok_rows[m] * z.K : z.A + (int64_t)m * z.lda) + k0;
              if (h0) u[i][0] = __ldcg(reinterpret_cast<const uint4*>(src + pc0 * 8));
              if (h1) u[i][1] = __ldcg(reinterpret_cast<const uint4*>(src + pc1 * 8));
            }
          }
          if (!rms) {
#pragma unroll
            for (int i = 0; i < RBATCH; ++i) {
              const int m = mb + i;
              if (m < rows) {
                if (h0) sts_v4(cx.act + m * cx.pitch + pc0 * 16, u[i][0]);
                if (h1) sts_v4(cx.act + m * cx.pitch + pc1 * 16, u[i][1]);
              }
            }
            continue;
          }
          // RMSNorm with HF rounding: y = w * bf16(x * rstd), rstd from fp32 sums over the whole row
#pragma unroll
          for (int i = 0; i < RBATCH; ++i) {
            float ss = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float2 a0 = unpack_bf16x2(u[i][h].x), a1 = unpack_bf16x2(u[i][h].y), a2 = unpack_bf16x2(u[i][h].z),
                           a3 = unpack_bf16x2(u[i][h].w);
              ss += a0.x * a0.x + a0.y * a0.y + a1.x * a1.x + a1.y * a1.y + a2.x * a2.x + a2.y * a2.y + a3.x * a3.x +
                    a3.y * a3.y;
            }
            ss = warp_sum(ss);
            if (lane == 0 && mb + i < rows) red[(mb + i) * RW + warp] = ss;
          }
          consumer_sync();
#pragma unroll
          for (int i = 0; i < RBATCH; ++i) {
            const int m = mb + i;
            if (m < rows) {
              float tot = 0.f;
#pragma unroll
              for (int w = 0; w < RW; ++w) tot += red[m * RW + w];
              const float rstd = rsqrtf(tot / (float)z.K + p.cfg.rms_eps);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h == 0 ? h0 : h1) {
                  const uint4 gw = h == 0 ? g0 : g1;
                  const uint32_t uu[4] = {u[i][h].x, u[i][h].y, u[i][h].z, u[i][h].w};
                  const uint32_t gg[4] = {gw.x, gw.y, gw.z, gw.w};
                  uint32_t oo[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    // two elements at a time: bf16(x * rstd) by one packed conversion, times the weight by one
                    // packed bf16 multiply (the product of two bf16 is exact in fp32, so rounding it once is what
                    // the fp32 form w * bf16_round(x * rstd) -> bf16 gives)
                    const float2 xv = unpack_bf16x2(uu[e]);
                    const uint32_t t = pack_bf16x2(xv.x * rstd, xv.y * rstd);
                    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(oo[e]) : "r"(gg[e]), "r"(t));
                  }
                  sts_v4(cx.act + m * cx.pitch + (h == 0 ? pc0 : pc1) * 16, make_uint4(oo[0], oo[1], oo[2], oo[3]));
                }
              }
            }
          }
        }
      }
    }
    consumer_sync();
    cx.stamp();

    // ---- consume the ring: this warp's span of the run ----
    const int C = b - a;
    const int w_lo = a + C * warp / RW, w_hi = a + C * (warp + 1) / RW;
    int np = 0, cur_rg = -1, cnt = 0;
    float acc[NT][4];
    // activation address of this lane for the B fragments: row (lane & 7) (+ 8 for lanes 16..31 when NT == 2), k half
    // (lane >> 3) & 1; rows beyond `rows` read row 0 (their products are never stored)
    int am = (lane & 7) + (NT == 2 ? (lane >> 4) * 8 : 0);
    if (am >= rows) am = 0;
    const uint32_t a_lane0 = cx.act + am * cx.pitch + ((lane >> 3) & 1) * 16;

    auto flush = [&]() {
      const bool whole = cnt == len;
      if (whole && g.KQ == 1 && !swiglu) {
        // the warp holds the complete sums of a row group: epilogue straight from the fragments
        const int wr = lane >> 2, mc = (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          store_out(z.out, z.ldo, z.epi, rows, n_out, cur_rg * TR + wr, nt * 8 + mc, acc[nt][0]);
          store_out(z.out, z.ldo, z.epi, rows, n_out, cur_rg * TR + wr, nt * 8 + mc + 1, acc[nt][1]);
          store_out(z.out, z.ldo, z.epi, rows, n_out, cur_rg * TR + wr + 8, nt * 8 + mc, acc[nt][2]);
          store_out(z.out, z.ldo, z.epi, rows, n_out, cur_rg * TR + wr + 8, nt * 8 + mc + 1, acc[nt][3]);
        }
      } else {
        if (np >= PW) __trap();  // the host-side sizing guarantees this never happens
        const int ti = warp * PW + np;
        ++np;
        float* tile = cx.pool + ti * TS;
        const int wr = lane >> 2, mc = (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          if (nt * 8 + mc < RS) {  // (RS is even: both rows of the pair exist or neither)
            *reinterpret_cast<float2*>(tile + wr * RS + nt * 8 + mc) = make_float2(acc[nt][0], acc[nt][1]);
            *reinterpret_cast<float2*>(tile + (wr + 8) * RS + nt * 8 + mc) = make_float2(acc[nt][2], acc[nt][3]);
          }
        }
        if (lane == 0) cx.prg[ti] = cur_rg;
      }
    };

    for (int cc = w_lo; cc < w_hi; ++cc) {
      const uint32_t e = cx.n_used++;
      const uint32_t lap = e / (uint32_t)cx.depth, slot = e - lap * (uint32_t)cx.depth;
      mbar_wait(cx.bars + 8u * slot, lap & 1u);
      const int rem = cc - qbase;
      const int rgi = rem / len, kl = rem - rgi * len;
      if (rgi != cur_rg) {
        if (cur_rg >= 0) flush();
        cur_rg = rgi;
        cnt = 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      }
      mma_tile<NT>(cx.ring + slot * SLOT_BYTES, a_lane0 + kl * (TK * 2), acc);
      ++cnt;
      __syncwarp();
      refill(p, cx, slot);  // the slot is free: request the tile this warp needs `depth` tiles from now
    }
    if (cur_rg >= 0) flush();
    consumer_sync();
    after_consume();
    cx.stamp();

    // ---- per output group: add the warps' pieces; finish here or through the global ticket ----
    const int rg_first = (a - qbase) / len, rg_last = (b - 1 - qbase) / len;
    const int og_first = swiglu ? rg_first >> 1 : rg_first, og_last = swiglu ? rg_last >> 1 : rg_last;
    for (int og = og_first + warp; og <= og_last; og += RW) {
      float sum[2][EPL];
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < EPL; ++i) sum[s][i] = 0.f;
      bool any = false;
      int contrib = 0;
      int seg_n[2] = {0, 0};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s == 1 && !swiglu) break;
        const int rgi = swiglu ? 2 * og + s : og;
        const int mine = lane < RW * PW ? cx.prg[lane] : -2;
        unsigned int mask = __ballot_sync(0xffffffffu, mine == rgi);
        while (mask) {
          const int ti = __ffs(mask) - 1;
          mask &= mask - 1;
          const float* tile = cx.pool + ti * TS + lane * eh;
#pragma unroll
          for (int i = 0; i < EPL; ++i)
            if (i < eh) sum[s][i] += tile[i];
          any = true;
        }
        const int s_lo = qbase + rgi * len;
        seg_n[s] = max(0, min(b, s_lo + len) - max(a, s_lo));
        contrib += seg_n[s];
      }
      if (!any) continue;  // every piece of the group was finished by the warp that held it
      const int wr = lane >> 1, m0 = (lane & 1) * eh;
      const int col = og * TR + wr;
      auto finish = [&]() {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          if (i >= eh) break;
          const int m = m0 + i;
          if (swiglu) {
            if (m < rows && col < n_out)
              reinterpret_cast<bf16*>(z.out)[(int64_t)m * z.ldo + col] = __float2bfloat16_rn(silu(sum[0][i]) * sum[1][i]);
          } else {
            store_out(z.out, z.ldo, z.epi, rows, n_out, col, m, sum[0][i]);
          }
        }
      };
      const int og_c0 = swiglu ? 2 * og * g.ck : og * g.ck;  // (KQ == 1) first chunk of the group
      if (g.KQ == 1 && og_c0 >= lo && og_c0 + (int)og_total <= hi) {
        finish();
        continue;
      }
      // pieces of other CTAs are missing: publish ours, the last contributor merges
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s == 1 && !swiglu) break;
        if (seg_n[s] == 0) continue;
        const int rgi = swiglu ? 2 * og + s : og;
        const int sidx = swiglu ? s : q;
        const int pi = bid - chunk_owner(qbase + rgi * len, g.T, nb);
        if (pi < 0 || pi >= p.maxcp) __trap();
        float* dst = p.pieces + ((int64_t)(og * nseg + sidx) * p.maxcp + pi) * TS + lane * eh;
#pragma unroll
        for (int i = 0; i < EPL; ++i)
          if (i < eh) dst[i] = sum[s][i];
      }
      __syncwarp();
      unsigned int old = 0;
      if (lane == 0) old = atom_release_add(p.tickets + og, (unsigned int)contrib);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old + (unsigned int)contrib != og_total) continue;
      // last contributor: add all pieces in a fixed order (k-part / gate-up, then CTA order); the loads of a batch of
      // pieces are issued together (one L2 round trip per batch, not per piece)
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < EPL; ++i) sum[s][i] = 0.f;
      constexpr int FB = 8;
      const int n_f = nseg * p.maxcp;
      for (int f0 = 0; f0 < n_f; f0 += FB) {
        float v[FB][EPL];
        int tgt[FB];
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const int f = f0 + j;
          tgt[j] = -1;
#pragma unroll
          for (int i = 0; i < EPL; ++i) v[j][i] = 0.f;
          if (f < n_f) {
            const int sidx = f / p.maxcp, pi = f - sidx * p.maxcp;
            const int qq = swiglu ? 0 : sidx;
            const int rgi = swiglu ? 2 * og + sidx : og;
            const int s_lo = seg_start(g, qq, rgi), s_len = part_len(g, qq);
            const int first = chunk_owner(s_lo, g.T, nb), last = chunk_owner(s_lo + s_len - 1, g.T, nb);
            bool valid = pi <= last - first;
            if (valid && g.T < nb) valid = range_lo(g.T, first + pi + 1, nb) > range_lo(g.T, first + pi, nb);  // CTA without chunks
            if (valid) {
              tgt[j] = swiglu ? sidx : 0;
              const float* src = p.pieces + ((int64_t)(og * nseg + sidx) * p.maxcp + pi) * TS + lane * eh;
#pragma unroll
              for (int i = 0; i < EPL; ++i)
                if (i < eh) v[j][i] = __ldcg(src + i);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          if (tgt[j] >= 0) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
              if (tgt[j] == 0) sum[0][i] += v[j][i];
              else sum[1][i] += v[j][i];
            }
          }
        }
      }
      finish();
      if (lane == 0) p.tickets[og] = 0;  // (the next phase that uses it starts after a grid barrier)
    }
    a = b;
    if (a < hi) consumer_sync();  // the next run restages the activations and reuses the pool
  }
}

// ---- attention ---------------------------------------------------------------------------------------------------
struct AttLayout {
  uint32_t q, k, v, pb;  // shared-memory addresses
  float* wmax;           // [RW][64]
  float* wsum;           // [RW][64]
  uint32_t* kbits;       // [SUBK] beams allowed to see each key
};

struct AttGeom {
  int t, g_cur, pos_cur, Nk, n_splits, n_items, n_inputs, QR, QT;
};
__device__ __forceinline__ AttGeom att_geom(const RowsParams& p, int t) {
  AttGeom a;
  a.t = t;
  a.g_cur = t - 1;
  a.pos_cur = p.S + t - 1;
  a.Nk = p.S + p.beams * t;
  a.n_splits = (a.Nk + SUBK - 1) / SUBK;
  a.n_inputs = p.rows / p.beams;
  a.n_items = a.n_inputs * p.cfg.n_kv_heads * a.n_splits;
  a.QR = (p.beams * GQ + 15) & ~15;
  a.QT = a.QR >> 4;
  return a;
}

// Request the K and V rows of an item that already exist in the caches (everything but this step's own rows) with
// one 256-byte bulk copy per row; rows past the end of the key list are zeroed (0 x garbage must not become NaN).
__device__ __forceinline__ void att_request(const RowsParams& p, const AttLayout& s, const AttGeom& ag, int layer,
                                            int item, uint32_t kvbar) {
  const int tid = threadIdx.x;
  if (tid >= KV_ISSUERS) return;
  const int KVH = p.cfg.n_kv_heads, kvd = KVH * HD;
  const int which = tid >> 6, j = tid & (SUBK - 1);
  const int pair = item / ag.n_splits, split = item - pair * ag.n_splits;
  const int input = pair / KVH, kvh = pair - input * KVH;
  const int key = split * SUBK + j;
  const int64_t n_prompt = (int64_t)ag.n_inputs * p.S, n_gen = (int64_t)p.rows * p.max_gen;
  const uint32_t dst = (which ? s.v : s.k) + j * KVP;
  const bf16* src = nullptr;
  bool zero = false;
  if (key < p.S) {
    src = p.kv_prompt + ((int64_t)layer * 2 + which) * n_prompt * kvd + ((int64_t)input * p.S + key) * kvd + kvh * HD;
  } else if (key < ag.Nk) {
    const int jj = key - p.S, gi = jj / p.beams, r = jj - gi * p.beams;
    if (gi < ag.g_cur)
      src = p.kv_gen + ((int64_t)layer * 2 + which) * n_gen * kvd +
            ((int64_t)(input * p.beams + r) * p.max_gen + gi) * kvd + kvh * HD;
    // gi == g_cur: this step's own row, written from qkv after the grid barrier
  } else {
    zero = true;
  }
  if (src != nullptr) {
    mbar_arrive_expect_tx(kvbar, 256);
    bulk_g2s(dst, src, 256, kvbar);
  } else {
    if (zero) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sts_v4(dst + i * 16, make_uint4(0, 0, 0, 0));
    }
    mbar_arrive(kvbar);
  }
}

// Beams allowed to see key `threadIdx.x` of an item (threads 0..63): prompt keys by the left-pad mask, generated
// (step, physical row) entries by the ancestry table - all 16 loads of it issued together.
__device__ __forceinline__ uint32_t att_key_bits(const RowsParams& p, const AttGeom& ag, int item, uint32_t all_beams) {
  const int tid = threadIdx.x;
  uint32_t bits = 0;
  if (tid >= SUBK) return bits;
  const int KVH = p.cfg.n_kv_heads;
  const int pair = item / ag.n_splits, split = item - pair * ag.n_splits;
  const int input = pair / KVH;
  const int key = split * SUBK + tid;
  if (key < p.S) {
    bits = (p.prompt_valid == nullptr || p.prompt_valid[(int64_t)input * p.S + key] != 0) ? all_beams : 0u;
  } else if (key < ag.Nk) {
    const int jj = key - p.S, gi = jj / p.beams, r = jj - gi * p.beams;
    if (gi == ag.g_cur) {
      bits = 1u << r;
    } else {
      int sl[16];
#pragma unroll
      for (int bb = 0; bb < 16; ++bb)
        sl[bb] = bb < p.beams ? p.slots[(int64_t)(input * p.beams + bb) * p.max_gen + gi] : -1;
#pragma unroll
      for (int bb = 0; bb < 16; ++bb)
        if (sl[bb] == input * p.beams + r) bits |= 1u << bb;
    }
  }
  return bits;
}

template <int NT>
__device__ __forceinline__ void attention_phase(const RowsParams& p, Ctx& cx, GridBarrier& bar, const AttLayout& s,
                                                uint32_t kvbar, uint32_t& kv_par, int layer, int t, bool requested) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = gridDim.x, bid = blockIdx.x;
  const int H = p.cfg.n_heads, KVH = p.cfg.n_kv_heads, kvd = KVH * HD, qkv_dim = (H + 2 * KVH) * HD;
  const AttGeom ag = att_geom(p, t);
  const int it_lo = range_lo(ag.n_items, bid, nb), it_hi = range_lo(ag.n_items, bid + 1, nb);
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int64_t n_gen = (int64_t)p.rows * p.max_gen;
  bf16* kg = p.kv_gen + ((int64_t)layer * 2 + 0) * n_gen * kvd;
  bf16* vg = p.kv_gen + ((int64_t)layer * 2 + 1) * n_gen * kvd;
  const float2* cs = reinterpret_cast<const float2*>(p.rope) + (int64_t)ag.pos_cur * (HD / 2);
  const uint32_t all_beams = (1u << p.beams) - 1u;

  // the first item's cached rows are requested before the grid barrier that completes qkv (normally right after the
  // qkv phase's last tile, see weight_phase's after_consume)
  if (!requested) {
    consumer_sync();  // everybody is done with the staging area of the qkv phase
    fence_proxy_async_smem();
    if (it_lo < it_hi) att_request(p, s, ag, layer, it_lo, kvbar);
  }
  // ... and so is everything else that does not depend on this step's qkv: the RoPE factors of this thread's column
  // pieces (pieces pp = tid & 7 and pp + 8 of every row it rotates) and the per-key beam masks of the first item
  float2 csr[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) csr[e] = cs[(tid & 7) * 8 + e];
  uint32_t bits_first = 0;
  if (it_lo < it_hi) bits_first = att_key_bits(p, ag, it_lo, all_beams);
  cx.stamp();
  bar.sync();
  cx.stamp();

  int cur_pair = -1;
  for (int item = it_lo; item < it_hi; ++item) {
    const int pair = item / ag.n_splits, split = item - pair * ag.n_splits;
    const int input = pair / KVH, kvh = pair - input * KVH;
    const int key0 = split * SUBK;
    if (item != it_lo) {
      consumer_sync();  // the previous item is done with the tiles
      fence_proxy_async_smem();
      att_request(p, s, ag, layer, item, kvbar);
    }
    // ---- Q (once per (input, kv head)), this step's own K / V rows that fall into the item, and the per-key beam
    // masks: every global load is issued before the first use (one L2 round trip for all three) ----
    // Q / K rows: work unit = (row, pair of 16-byte pieces pp and pp + 8 = the two halves rotate-half RoPE mixes)
    const bool new_q = pair != cur_pair;
    cur_pair = pair;
    const int n_live = p.beams * GQ;
    const int q_units = new_q ? ag.QR * 8 : 0;
    const int kcur0 = p.S + ag.g_cur * p.beams - key0;  // item-relative index of beam 0's own key
    const bool has_cur = kcur0 + p.beams > 0 && kcur0 < SUBK;
    const int k_units = has_cur ? p.beams * 8 : 0;       // own K rows (RoPE)
    const int v_units = has_cur ? p.beams * 16 : 0;      // own V rows (copy), one piece per unit
    constexpr int UPT = 4;  // units per thread: at most 64 * 8 + 16 * 8 + 16 * 16 = 896 <= 4 * 256
    uint4 ua[UPT], ub[UPT];
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      const int un = tid + i * RT;
      ua[i] = ub[i] = make_uint4(0, 0, 0, 0);
      if (un < q_units) {
        const int r = un >> 3, pp = un & 7;
        if (r < n_live) {
          const bf16* src = p.qkv + (int64_t)(input * p.beams + r / GQ) * qkv_dim + (kvh * GQ + r % GQ) * HD;
          ua[i] = __ldcg(reinterpret_cast<const uint4*>(src + pp * 8));
          ub[i] = __ldcg(reinterpret_cast<const uint4*>(src + HD / 2 + pp * 8));
        }
      } else if (un < q_units + k_units) {
        const int r = (un - q_units) >> 3, pp = (un - q_units) & 7;
        const bf16* src = p.qkv + (int64_t)(input * p.beams + r) * qkv_dim + (H + kvh) * HD;
        ua[i] = __ldcg(reinterpret_cast<const uint4*>(src + pp * 8));
        ub[i] = __ldcg(reinterpret_cast<const uint4*>(src + HD / 2 + pp * 8));
      } else if (un < q_units + k_units + v_units) {
        const int r = (un - q_units - k_units) >> 4, pc = (un - q_units - k_units) & 15;
        ua[i] = __ldcg(reinterpret_cast<const uint4*>(p.qkv + (int64_t)(input * p.beams + r) * qkv_dim +
                                                       (H + KVH + kvh) * HD + pc * 8));
      }
    }
    // which beams may see each key (threads 0..63, one key each; the first item's were fetched before the barrier)
    const uint32_t bits = item == it_lo ? bits_first : att_key_bits(p, ag, item, all_beams);
    // rotate-half RoPE of a pair of pieces: out_lo = lo c - hi s, out_hi = hi c + lo s, rounded to bf16
    auto rope_pair = [&](const uint4& lo4, const uint4& hi4, int pp, uint4& o_lo, uint4& o_hi) {
      const uint32_t l[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, h[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
      uint32_t ol[4], oh[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 lv = unpack_bf16x2(l[e]), hv = unpack_bf16x2(h[e]);
        const float2 c0 = csr[e * 2], c1 = csr[e * 2 + 1];  // (pp == tid & 7 for every unit of this thread)
        ol[e] = pack_bf16x2(lv.x * c0.x - hv.x * c0.y, lv.y * c1.x - hv.y * c1.y);
        oh[e] = pack_bf16x2(hv.x * c0.x + lv.x * c0.y, hv.y * c1.x + lv.y * c1.y);
      }
      o_lo = make_uint4(ol[0], ol[1], ol[2], ol[3]);
      o_hi = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    };
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      const int un = tid + i * RT;
      if (un < q_units) {
        const int r = un >> 3, pp = un & 7;
        uint4 o_lo, o_hi;
        rope_pair(ua[i], ub[i], pp, o_lo, o_hi);  // (zero rows stay zero)
        sts_v4(s.q + r * KVP + pp * 16, o_lo);
        sts_v4(s.q + r * KVP + (pp + 8) * 16, o_hi);
      } else if (un < q_units + k_units) {
        const int r = (un - q_units) >> 3, pp = (un - q_units) & 7;
        const int kj = kcur0 + r;
        if (kj >= 0 && kj < SUBK) {
          uint4 o_lo, o_hi;
          rope_pair(ua[i], ub[i], pp, o_lo, o_hi);
          sts_v4(s.k + kj * KVP + pp * 16, o_lo);
          sts_v4(s.k + kj * KVP + (pp + 8) * 16, o_hi);
          bf16* dstk = kg + ((int64_t)(input * p.beams + r) * p.max_gen + ag.g_cur) * kvd + kvh * HD;
          *reinterpret_cast<uint4*>(dstk + pp * 8) = o_lo;
          *reinterpret_cast<uint4*>(dstk + HD / 2 + pp * 8) = o_hi;
        }
      } else if (un < q_units + k_units + v_units) {
        const int r = (un - q_units - k_units) >> 4, pc = (un - q_units - k_units) & 15;
        const int kj = kcur0 + r;
        if (kj >= 0 && kj < SUBK) {
          sts_v4(s.v + kj * KVP + pc * 16, ua[i]);
          bf16* dstv = vg + ((int64_t)(input * p.beams + r) * p.max_gen + ag.g_cur) * kvd + kvh * HD;
          *reinterpret_cast<uint4*>(dstv + pc * 8) = ua[i];
        }
      }
    }
    if (tid < SUBK) s.kbits[tid] = bits;
    consumer_sync();
    if (item == it_lo) cx.stamp();
    mbar_wait(kvbar, kv_par);
    kv_par ^= 1u;
    if (item == it_lo) cx.stamp();

    // ---- S = Q K^T: warp w owns keys 8 w .. 8 w + 7 ----
    float sc[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) sc[mt][0] = sc[mt][1] = sc[mt][2] = sc[mt][3] = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < HD / 32; ++k2) {  // two k-steps of 16 dims per iteration
      uint32_t kb[4];                        // (k-step 0 lo, hi), (k-step 1 lo, hi) for the warp's 8 keys
      ldsm_x4(s.k + (warp * 8 + (lane & 7)) * KVP + (k2 * 4 + (lane >> 3)) * 16, kb);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < ag.QT) {
          const uint32_t qa = s.q + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * KVP + (lane >> 4) * 16;
          uint32_t af[4];
          ldsm_x4(qa + (k2 * 4) * 16, af);
          mma16816(sc[mt], af, kb[0], kb[1]);
          ldsm_x4(qa + (k2 * 4 + 2) * 16, af);
          mma16816(sc[mt], af, kb[2], kb[3]);
        }
      }
    }
    // mask + scale; per-warp row maxima
    const int kq = warp * 8 + (lane & 3) * 2;  // this lane's two keys
    const uint32_t bits0 = s.kbits[kq], bits1 = s.kbits[kq + 1];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt < ag.QT) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int qr = mt * 16 + (lane >> 2) + hf * 8;
          const int bm = qr / GQ;
          float v0 = ((bits0 >> bm) & 1u) ? sc[mt][hf * 2] * scale_log2 : -INFINITY;
          float v1 = ((bits1 >> bm) & 1u) ? sc[mt][hf * 2 + 1] * scale_log2 : -INFINITY;
          sc[mt][hf * 2] = v0;
          sc[mt][hf * 2 + 1] = v1;
          float mx = fmaxf(v0, v1);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          if ((lane & 3) == 0) s.wmax[warp * 64 + qr] = mx;
        }
      }
    }
    consumer_sync();
    // probabilities (bf16 for the P.V MMA, fp32 for the row sums)
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt < ag.QT) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int qr = mt * 16 + (lane >> 2) + hf * 8;
          float mx = -INFINITY;
#pragma unroll
          for (int w = 0; w < RW; ++w) mx = fmaxf(mx, s.wmax[w * 64 + qr]);
          const float p0 = (mx == -INFINITY) ? 0.f : exp2f(sc[mt][hf * 2] - mx);
          const float p1 = (mx == -INFINITY) ? 0.f : exp2f(sc[mt][hf * 2 + 1] - mx);
          sts_u32(s.pb + qr * PPB + kq * 2, pack_bf16x2(p0, p1));
          float sm = p0 + p1;
          sm += __shfl_xor_sync(0xffffffffu, sm, 1);
          sm += __shfl_xor_sync(0xffffffffu, sm, 2);
          if ((lane & 3) == 0) s.wsum[warp * 64 + qr] = sm;
        }
      }
    }
    consumer_sync();
    if (item == it_lo) cx.stamp();
    // ---- O = P V: warp w owns output dims 16 w .. 16 w + 15 ----
    float oc[4][2][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int n = 0; n < 2; ++n) oc[mt][n][0] = oc[mt][n][1] = oc[mt][n][2] = oc[mt][n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < SUBK / 16; ++ks) {
      uint32_t vb[4];  // (keys lo, dims lo) (keys hi, dims lo) (keys lo, dims hi) (keys hi, dims hi)
      ldsm_x4_t(s.v + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * KVP + (warp * 2 + (lane >> 4)) * 16, vb);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < ag.QT) {
          uint32_t af[4];
          ldsm_x4(s.pb + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PPB + (ks * 2 + (lane >> 4)) * 16, af);
          mma16816(oc[mt][0], af, vb[0], vb[1]);
          mma16816(oc[mt][1], af, vb[2], vb[3]);
        }
      }
    }
    // partials: [row][KVH][max_splits][GQ][PSTR]
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt < ag.QT) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int qr = mt * 16 + (lane >> 2) + hf * 8;
          if (qr < p.beams * GQ) {
            const int row = input * p.beams + qr / GQ, hh = qr % GQ;
            float* part = p.part + ((((int64_t)row * KVH + kvh) * p.max_splits + split) * GQ + hh) * PSTR;
#pragma unroll
            for (int n = 0; n < 2; ++n)
              *reinterpret_cast<float2*>(part + warp * 16 + n * 8 + (lane & 3) * 2) =
                  make_float2(oc[mt][n][hf * 2], oc[mt][n][hf * 2 + 1]);
          }
        }
      }
    }
    if (tid < p.beams * GQ) {
      const int row = input * p.beams + tid / GQ, hh = tid % GQ;
      float* part = p.part + ((((int64_t)row * KVH + kvh) * p.max_splits + split) * GQ + hh) * PSTR;
      float mx = -INFINITY, sm = 0.f;
#pragma unroll
      for (int w = 0; w < RW; ++w) {
        mx = fmaxf(mx, s.wmax[w * 64 + tid]);
        sm += s.wsum[w * 64 + tid];
      }
      part[HD] = mx;
      part[HD + 1] = sm;
    }
  }
  if (it_lo >= it_hi) { cx.stamp(); cx.stamp(); cx.stamp(); }  // (same number of stamps with and without items)
  cx.stamp();
  bar.sync();  // every partial of the layer is in L2
  cx.stamp();

  // ---- merge: one warp per (row, head), lane = 4 output dims ----
  // Lane s (and s + 32) fetches the (max, sum) pair of split s while every lane already fetches its 4 output dims of
  // the first MERGE_B splits: one L2 round trip for up to MERGE_B splits, one more per further MERGE_B.
  const int n_units = p.rows * H;
  for (int u = bid + nb * warp; u < n_units; u += nb * RW) {
    const int row = u / H, h = u - row * H;
    const int kvh = h / GQ, hh = h - kvh * GQ;
    const float* base = p.part + ((((int64_t)row * KVH + kvh) * p.max_splits) * GQ + hh) * PSTR;
    const int64_t stride = (int64_t)GQ * PSTR;
    float l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, mx = -INFINITY;
    for (int sb = 0; sb < ag.n_splits; sb += 64) {  // (more than 64 splits: running rescale between blocks)
      const int nsb = min(64, ag.n_splits - sb);
      const float2 ml0 = lane < nsb ? __ldcg(reinterpret_cast<const float2*>(base + (sb + lane) * stride + HD))
                                    : make_float2(-INFINITY, 0.f);
      const float2 ml1 = lane + 32 < nsb ? __ldcg(reinterpret_cast<const float2*>(base + (sb + lane + 32) * stride + HD))
                                         : make_float2(-INFINITY, 0.f);
      float4 vv[MERGE_B];
#pragma unroll
      for (int i = 0; i < MERGE_B; ++i)
        vv[i] = i < nsb ? __ldcg(reinterpret_cast<const float4*>(base + (sb + i) * stride + lane * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      const float bm = warp_max(fmaxf(ml0.x, ml1.x));
      const float nm = fmaxf(mx, bm);
      const float resc = (mx == -INFINITY) ? 0.f : exp2f(mx - nm);
      l *= resc; a0 *= resc; a1 *= resc; a2 *= resc; a3 *= resc;
      mx = nm;
      const float w0 = (ml0.x == -INFINITY) ? 0.f : exp2f(ml0.x - mx), w1 = (ml1.x == -INFINITY) ? 0.f : exp2f(ml1.x - mx);
      l += warp_sum(w0 * ml0.y + w1 * ml1.y);
      for (int s0 = 0; s0 < nsb; s0 += MERGE_B) {
        if (s0 > 0) {
#pragma unroll
          for (int i = 0; i < MERGE_B; ++i)
            vv[i] = s0 + i < nsb ? __ldcg(reinterpret_cast<const float4*>(base + (sb + s0 + i) * stride + lane * 4))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < MERGE_B; ++i) {
          const int si = s0 + i;  // weight of split si lives in lane si & 31 (w0 for si < 32, w1 otherwise)
          const float wa = __shfl_sync(0xffffffffu, w0, si & 31), wb = __shfl_sync(0xffffffffu, w1, si & 31);
          const float w = si < 32 ? wa : wb;
          a0 = fmaf(w, vv[i].x, a0);
          a1 = fmaf(w, vv[i].y, a1);
          a2 = fmaf(w, vv[i].z, a2);
          a3 = fmaf(w, vv[i].w, a3);
        }
      }
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    uint2 o;
    o.x = pack_bf16x2(a0 * inv, a1 * inv);
    o.y = pack_bf16x2(a2 * inv, a3 * inv);
    *reinterpret_cast<uint2*>(p.attn + (int64_t)row * (H * HD) + h * HD + lane * 4) = o;
  }
}

template <int NT>
__global__ void __launch_bounds__(RBLOCK, 1)
llama_decode_rows_megakernel(const RowsParams p) {
  extern __shared__ uint8_t rows_smem_raw[];
  const pcy_llama_config& c = p.cfg;
  const int d = c.d_model;
  const int ns = p.ring_slots;
  // layout (1 KB aligned): [ring: ns slots][work area: activations | attention tiles][pool][misc: pool row groups,
  // tokens, mbarriers, RMSNorm weight pointers]
  const uint32_t raw = smem_u32(rows_smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* base = rows_smem_raw + pad;
  const uint32_t ring = raw + pad;
  uint8_t* work = base + (size_t)ns * SLOT_BYTES;
  uint8_t* misc = work + p.off_misc;
  const uint32_t bars = smem_u32(misc + MISC_BARS);
  const uint32_t kvbar = bars + 8u * ns;
  // RMSNorm weight pointers of every layer, copied once so that no phase starts with a dependent global load
  const bf16** s_ln = reinterpret_cast<const bf16**>(misc + MISC_LN);
  for (int i = threadIdx.x; i < 2 * c.n_layers; i += RBLOCK)
    s_ln[i] = (i & 1) ? p.layers[i >> 1].ln2 : p.layers[i >> 1].ln1;
  // this CTA's chunk range of every phase kind (the cut itself: cta_lo)
  int* s_lohi = reinterpret_cast<int*>(misc + MISC_LOHI);
  if (threadIdx.x < 5) {
    s_lohi[2 * threadIdx.x] = cta_lo(p.geom[threadIdx.x], blockIdx.x, gridDim.x);
    s_lohi[2 * threadIdx.x + 1] = cta_lo(p.geom[threadIdx.x], blockIdx.x + 1, gridDim.x);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < ns; ++i) mbar_init(bars + 8u * i, 1);  // one arrive.expect_tx by the refilling warp + the bytes
    mbar_init(kvbar, KV_ISSUERS);
    fence_barrier_init();
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int n_phases = 4 * c.n_layers + 1;
  Ctx cx;
  {
    const int warp = tid >> 5;
    cx.depth = ns / RW;
    cx.ring = ring + (uint32_t)(warp * cx.depth) * SLOT_BYTES;
    cx.bars = bars + 8u * (uint32_t)(warp * cx.depth);
    cx.n_used = 0;
    cx.n_phases = n_phases;
    cx.fill.ph = 0;
    cx.fill.lohi = s_lohi;
    cx.fill.begin_phase(p, warp);
    for (int i = 0; i < cx.depth; ++i) refill(p, cx, (uint32_t)i);  // the warp's first tiles
  }
  cx.act = smem_u32(work);
  cx.pitch = (uint32_t)p.kcap * 2u + 16u;
  cx.pool = reinterpret_cast<float*>(work + p.off_pool);
  cx.rs = (p.rows + 1) & ~1;
  cx.prg = reinterpret_cast<int*>(misc);
  cx.tbuf = p.timing;
  cx.tix = 0;
  cx.tcta = p.timing_cta;
  AttLayout al;
  {
    const int QR = (p.beams * GQ + 15) & ~15;
    al.q = smem_u32(work);
    al.k = al.q + QR * KVP;
    al.v = al.k + SUBK * KVP;
    al.pb = al.v + SUBK * KVP;
    uint8_t* after = work + (size_t)QR * KVP + 2 * SUBK * KVP + (size_t)QR * PPB;
    al.wmax = reinterpret_cast<float*>(after);
    al.wsum = al.wmax + RW * 64;
    al.kbits = reinterpret_cast<uint32_t*>(al.wsum + RW * 64);
  }
  int32_t* s_tok = reinterpret_cast<int32_t*>(misc + MISC_TOK);  // [rows] token of every row (layer 0 reads the table)

  GridBarrier bar{p.barrier, 0u, gridDim.x};
  const int t = p.state[0];
  uint32_t kv_par = 0;
  cx.stamp();

  // the residual stream starts as the embedding of the last token of every row: each CTA mirrors its column slice
  if (tid < p.rows) s_tok[tid] = p.tokens[(int64_t)tid * p.max_gen + (t - 1)];
  consumer_sync();
  {
    const int n8 = d / 8;
    const int clo = range_lo(n8, blockIdx.x, gridDim.x), chi = range_lo(n8, blockIdx.x + 1, gridDim.x);
    for (int m = 0; m < p.rows; ++m) {
      const bf16* src = p.embed + (int64_t)s_tok[m] * d;
      for (int k8 = clo + tid; k8 < chi; k8 += RT)
        *reinterpret_cast<uint4*>(p.x + (int64_t)m * d + k8 * 8) = *reinterpret_cast<const uint4*>(src + k8 * 8);
    }
  }

  bool kv_requested = false;
  for (int ph = 0; ph < n_phases; ++ph) {
    const int kind = ph == n_phases - 1 ? 4 : (ph & 3);
    const PhaseDesc z = phase_desc(p, s_ln, ph);
    if (kind == 1) {
      attention_phase<NT>(p, cx, bar, al, kvbar, kv_par, ph >> 2, t, kv_requested);
      kv_requested = false;
      cx.stamp();
      bar.sync();
      cx.stamp();
    }
    weight_phase<NT>(p, cx, z, ph == 0 ? s_tok : nullptr, [&]() {
      if (kind != 0) return;
      kv_requested = true;  // (a CTA without qkv tiles never gets here: attention_phase then asks itself)
      // the attention phase of this layer follows: the staging area is free, ask for the K / V rows of this CTA's
      // first item now - they travel while the qkv epilogue and the grid barrier run
      const AttGeom ag = att_geom(p, t);
      const int it_lo = range_lo(ag.n_items, blockIdx.x, gridDim.x), it_hi = range_lo(ag.n_items, blockIdx.x + 1, gridDim.x);
      fence_proxy_async_smem();
      if (it_lo < it_hi) att_request(p, al, ag, ph >> 2, it_lo, kvbar);
    });
    cx.stamp();
    if (kind != 0 && kind != 4) {  // (the barrier after the qkv phase is inside attention_phase)
      bar.sync();
      cx.stamp();
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
constexpr size_t SMEM_LIMIT = 227 * 1024;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn rows_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  });
  return fn;
}

struct RowsPlan {
  int nt, kcap, ring_slots, maxcp, off_pool, off_misc;
  size_t smem;
  int64_t max_og, piece_floats;
  bool ok;
};

// everything the kernel's bookkeeping needs to hold for one matrix: pool tiles per warp, piece slots per segment
void plan_matrix(int N, int K, int kcap, bool swiglu, int nb, RowsPlan& pl) {
  const Geom g = make_geom(N, K, kcap);
  // chunks of the largest CTA range: stream-K cut, or (single-part matrices, tuning knob) cut at group boundaries
  int cmax = (g.T + nb - 1) / nb;
  if (g.KQ == 1) cmax = std::max(cmax, (g.n_rg + nb - 1) / nb * g.ck);
  const int span = (cmax + RW - 1) / RW;  // chunks of the largest warp span
  int len_min = g.ckq;
  for (int q = 0; q < g.KQ; ++q) len_min = std::min(len_min, part_len(g, q));
  // a span of s chunks touches at most (s + len - 2) / len + 1 row groups; complete ones of a plain single-part
  // matrix never reach the pool
  int pieces = span <= 1 ? 1 : (span + len_min - 2) / len_min + 1;
  if (g.KQ == 1 && !swiglu) pieces = std::min(pieces, 2);
  if (pieces > PW) pl.ok = false;
  if ((int64_t)g.T * (nb + 1) >= (1ll << 31)) pl.ok = false;  // 32-bit range arithmetic in the kernel
  // CTAs (empty ranges included) between the owners of the first and the last chunk of a segment
  for (int q = 0; q < g.KQ; ++q) {
    const int len = part_len(g, q);
    for (int rg = 0; rg < g.n_rg; ++rg) {
      const int s_lo = seg_start(g, q, rg);
      pl.maxcp = std::max(pl.maxcp, chunk_owner(s_lo + len - 1, g.T, nb) - chunk_owner(s_lo, g.T, nb) + 1);
    }
  }
  const int64_t n_og = swiglu ? g.n_rg / 2 : g.n_rg;
  const int64_t nseg = swiglu ? 2 : g.KQ;
  pl.max_og = std::max(pl.max_og, n_og);
  pl.piece_floats = std::max(pl.piece_floats, n_og * nseg);  // (x maxcp x tile floats, applied at the end)
}

RowsPlan compute_plan(const pcy_llama_config& c, int rows, int beams) {
  RowsPlan pl{};
  pl.ok = true;
  const int nb = num_sms();
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
  if (c.head_dim != HD || KVH <= 0 || H != GQ * KVH) pl.ok = false;
  if (rows < 1 || rows > 16 || beams < 1 || beams > 16 || rows % beams != 0) pl.ok = false;
  if (d % TK != 0 || f % TK != 0 || (H * HD) % TK != 0 || f % 2 != 0) pl.ok = false;
  if (d > 2 * RT * 8) pl.ok = false;  // the RMSNorm staging holds a whole row in two 16-byte pieces per thread
  if (!pl.ok) return pl;
  pl.nt = rows <= 8 ? 1 : 2;
  pl.kcap = std::max(d, H * HD);
  const int qkv_dim = (H + 2 * KVH) * HD;
  plan_matrix(qkv_dim, d, pl.kcap, false, nb, pl);
  plan_matrix(d, H * HD, pl.kcap, false, nb, pl);
  plan_matrix(2 * f, d, pl.kcap, true, nb, pl);
  plan_matrix(d, f, pl.kcap, false, nb, pl);
  plan_matrix(c.vocab, d, pl.kcap, false, nb, pl);
  const int ts = TR * ((rows + 1) & ~1);  // floats of a partial tile
  pl.piece_floats *= (int64_t)pl.maxcp * ts;
  // shared memory: work area = max(activations, attention tiles) | pool | misc (pool row groups, tokens, layer table)
  const int QR = (beams * GQ + 15) & ~15;
  const size_t act = (size_t)rows * (pl.kcap * 2 + 16);
  const size_t att = (size_t)QR * KVP + 2 * SUBK * KVP + (size_t)QR * PPB + 2 * RW * 64 * 4 + SUBK * 4;
  pl.off_pool = (int)round_up((int64_t)std::max(act, att), 128);
  const size_t pool = (size_t)RW * PW * ts * 4;
  pl.off_misc = (int)round_up(pl.off_pool + (int64_t)pool, 128);
  const size_t misc = MISC_LN + (size_t)c.n_layers * 16;
  const size_t fixed = 1024 /*alignment*/ + pl.off_misc + misc;
  if (fixed + RW * SLOT_BYTES > SMEM_LIMIT) {
    pl.ok = false;
    return pl;
  }
  int ns = (int)((SMEM_LIMIT - fixed) / SLOT_BYTES) / RW * RW;
  ns = std::min(ns, MAX_SLOTS);
  pl.ring_slots = ns;
  pl.smem = fixed + (size_t)ns * SLOT_BYTES;
  return pl;
}

// (the plan walks every segment of the LM head: computed once per (model shape, rows, beams))
RowsPlan make_plan(const pcy_llama_config& c, int rows, int beams) {
  struct Entry {
    pcy_llama_config c;
    int rows, beams, sms;
    RowsPlan pl;
  };
  static std::mutex mu;
  static std::vector<Entry> cache;
  const int sms = num_sms();
  std::lock_guard<std::mutex> lock(mu);
  for (const Entry& e : cache)
    if (e.rows == rows && e.beams == beams && e.sms == sms && e.c.n_layers == c.n_layers && e.c.d_model == c.d_model &&
        e.c.n_heads == c.n_heads && e.c.n_kv_heads == c.n_kv_heads && e.c.head_dim == c.head_dim &&
        e.c.ffn_dim == c.ffn_dim && e.c.vocab == c.vocab)
      return e.pl;
  if (cache.size() > 256) cache.clear();
  cache.push_back(Entry{c, rows, beams, sms, compute_plan(c, rows, beams)});
  return cache.back().pl;
}

int max_splits_for(int S, int beams, int max_gen) { return ceil_div(S + beams * max_gen, SUBK); }

unsigned long long* g_rows_timing = nullptr;

}  // namespace

void decode_rows_megakernel_set_timing(unsigned long long* dev_buf) { g_rows_timing = dev_buf; }

// Tensor maps of all weight matrices (2-D bf16, box 64 k x 16 rows, 128-byte swizzle, zero fill past the last row):
// host array of 4 L + 1 maps, uploaded by the caller.  Returns non-zero when the driver entry point is missing.
int decode_rows_build_maps(const pcy_llama_config& c, const LlamaLayerPtrs* layers_host, const bf16* lm_head,
                           void** maps_dev) {
  *maps_dev = nullptr;
  EncodeTiledFn enc = rows_encode_fn();
  if (!enc) return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
  if (d % 8 != 0 || f % 8 != 0) return set_error(PCY_ERR_UNSUPPORTED, "decode rows maps: unaligned rows");
  std::vector<CUtensorMap> maps(4 * c.n_layers + 1);
  auto encode = [&](CUtensorMap* out, const bf16* ptr, int64_t n_rows, int64_t cols) -> int {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)n_rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)TR};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PCY_ERR_CUDA, "cuTensorMapEncodeTiled(decode weights) failed (%d)", (int)r);
    return 0;
  };
  for (int l = 0; l < c.n_layers; ++l) {
    const LlamaLayerPtrs& y = layers_host[l];
    PCY_TRY(encode(&maps[4 * l + 0], y.wqkv, (int64_t)(H + 2 * KVH) * HD, d));
    PCY_TRY(encode(&maps[4 * l + 1], y.wo, d, (int64_t)H * HD));
    PCY_TRY(encode(&maps[4 * l + 2], y.wgu, 2 * (int64_t)f, d));
    PCY_TRY(encode(&maps[4 * l + 3], y.wdown, d, f));
  }
  PCY_TRY(encode(&maps[4 * c.n_layers], lm_head, c.vocab, d));
  void* dev = nullptr;
  PCY_CUDA(cudaMalloc(&dev, maps.size() * sizeof(CUtensorMap)));
  PCY_CUDA(cudaMemcpy(dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  *maps_dev = dev;
  return 0;
}

bool decode_rows_megakernel_supported(const pcy_llama_config& c, int rows, int beams) {
  const RowsPlan pl = make_plan(c, rows, beams);
  return pl.ok && pl.ring_slots >= RW;
}

int64_t decode_rows_megakernel_scratch_bytes(const pcy_llama_config& c, int rows, int beams, int S, int max_gen) {
  const RowsPlan pl = make_plan(c, rows, beams);
  if (!pl.ok) return 0;
  const int64_t d = c.d_model, qkv = (int64_t)(c.n_heads + 2 * c.n_kv_heads) * HD;
  int64_t b = 512;
  b += round_up(pl.max_og * 4, 256);
  b += round_up(rows * d * 2, 256) + round_up(rows * qkv * 2, 256) + round_up((int64_t)rows * c.n_heads * HD * 2, 256) +
       round_up((int64_t)rows * c.ffn_dim * 2, 256);
  b += round_up((int64_t)rows * c.n_kv_heads * max_splits_for(S, beams, max_gen) * GQ * PSTR * 4, 256);
  b += round_up(pl.piece_floats * 4, 256);
  return b + 1024;
}

int decode_rows_megakernel(const pcy_llama_config& c, const LlamaLayerPtrs* layers_dev, const void* maps_dev,
                           const bf16* embed, const bf16* norm, const float* rope, const pcy_decode_buffers* b,
                           void* scratch, cudaStream_t stream) {
  const int rows = b->n_inputs * b->beams;
  const RowsPlan pl = make_plan(c, rows, b->beams);
  PCY_REQUIRE(pl.ok && pl.ring_slots >= RW && maps_dev != nullptr, "decode rows megakernel: unsupported configuration");
  RowsParams p;
  p.cfg = c; p.embed = embed; p.norm = norm; p.layers = layers_dev;
  p.maps = reinterpret_cast<const CUtensorMap*>(maps_dev); p.rope = rope;
  p.rows = rows; p.beams = b->beams; p.S = b->S; p.max_gen = b->max_gen;
  p.kv_prompt = reinterpret_cast<const bf16*>(b->kv_prompt); p.prompt_valid = b->prompt_valid;
  p.kv_gen = reinterpret_cast<bf16*>(b->kv_gen); p.tokens = b->tokens; p.slots = b->slots; p.state = b->state;
  p.logits = b->logits_cur;
  const int64_t d = c.d_model, qkv = (int64_t)(c.n_heads + 2 * c.n_kv_heads) * HD;
  uint8_t* s = reinterpret_cast<uint8_t*>(round_up(reinterpret_cast<int64_t>(scratch), 256));
  auto carve = [&](int64_t bytes) { uint8_t* r = s; s += round_up(bytes, 256); return r; };
  p.barrier = reinterpret_cast<unsigned int*>(carve(512));
  p.tickets = reinterpret_cast<unsigned int*>(carve(pl.max_og * 4));
  p.x = reinterpret_cast<bf16*>(carve(rows * d * 2));
  p.qkv = reinterpret_cast<bf16*>(carve(rows * qkv * 2));
  p.attn = reinterpret_cast<bf16*>(carve((int64_t)rows * c.n_heads * HD * 2));
  p.act = reinterpret_cast<bf16*>(carve((int64_t)rows * c.ffn_dim * 2));
  p.max_splits = max_splits_for(b->S, b->beams, b->max_gen);
  p.part = reinterpret_cast<float*>(carve((int64_t)rows * c.n_kv_heads * p.max_splits * GQ * PSTR * 4));
  p.pieces = reinterpret_cast<float*>(s);
  p.ring_slots = pl.ring_slots;
  p.kcap = pl.kcap; p.maxcp = pl.maxcp; p.off_pool = pl.off_pool; p.off_misc = pl.off_misc;
  p.timing = g_rows_timing;
  static const int timing_cta = [] {
    const char* e = getenv("PCY_ROWS_TIMING_CTA");  // which CTA writes the profiling stamps (default 0)
    return e ? atoi(e) : 0;
  }();
  p.timing_cta = std::min(std::max(timing_cta, 0), num_sms() - 1);
  static const int align_env = [] {
    const char* e = getenv("PCY_ROWS_ALIGN");  // tuning knob, see RowsParams::align_mask
    return e ? atoi(e) : 3;
  }();
  p.align_mask = align_env;
  {
    const int d = c.d_model, f = c.ffn_dim, H = c.n_heads, KVH = c.n_kv_heads;
    const int Ns[5] = {(H + 2 * KVH) * HD, d, 2 * f, d, c.vocab}, Ks[5] = {d, H * HD, d, f, d};
    for (int k = 0; k < 5; ++k) p.geom[k] = make_geom(Ns[k], Ks[k], pl.kcap, (p.align_mask >> k) & 1);
  }
  PCY_CUDA(cudaMemsetAsync(p.barrier, 0, 512, stream));  // grid barrier counter (tickets reset themselves)
  void* fn = pl.nt == 1 ? (void*)llama_decode_rows_megakernel<1> : (void*)llama_decode_rows_megakernel<2>;
  static SmemOptIn opt[2];
  if (opt[pl.nt - 1].need(pl.smem))
    PCY_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  void* args[] = {(void*)&p};
  PCY_CUDA(cudaLaunchCooperativeKernel(fn, dim3(num_sms()), dim3(RBLOCK), args, pl.smem, stream));
  count_launch();
  return 0;
}

}  // namespace pcy
