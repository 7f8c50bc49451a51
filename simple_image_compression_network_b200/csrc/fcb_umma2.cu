// fcb_umma2.cu -- "umma_i8" main loop with a shared-memory-RESIDENT input patch, pixels on the MMA N axis.
//
// Why not the first-generation kernel (fcb_umma.cu)?  It re-fetches a 128-pixel A tile from L2 for every filter
// tap (each input byte crosses L2->SM ~6x for a 5x5 layer) and issues M=128 x N=128 x K=32 instructions, whose
// ~100-clock per-instruction floor (profiles/r01_umma_issue_pattern_study.log) caps the tensor pipe near 50 %.
//
// Here (1) the patch of input pixels a tile of R x WT output pixels needs is loaded ONCE per tile into "planes":
//
//   plane(py,px,cc)[j][i][0..127] = 128 channels (chunk cc) of input pixel (y, x) with
//        stride 2:  y = 2*(y0 + dy + j) + py,  x = 2*(x0 + dx + i) + px      (parity split of the NHWC image)
//        stride 1:  y = y0 + dy + j,           x = x0 + dx + i               (also the 4 phases of deconv522)
//   rows are P = WT + halo pixels wide and stored back to back, so output pixel (r, xo) of the tile, linear index
//   m = r*P + xo, reads for tap (offy, offx) plane row  m + (offy-dy)*P + (offx-dx): ONE contiguous, uniformly
//   strided K-major SWIZZLE_128B operand per tap -- the sliding window (slidingwindow.h:1254-1353) is a descriptor
//   offset.  Columns xo >= WT of each row are computed and discarded (WT/P efficiency).  Operand start addresses
//   that are multiples of 128 B (not 1024 B) are legal: the swizzle XOR uses absolute smem address bits, verified
//   on hardware (tools/umma_probe.cu, profiles/r01_umma_probe.log "P3 ... base_offset=0 : PASS").
//
// and (2) the GEMM is transposed:  D[ch][pixel] = W[ch][k] * Patch[pixel][k]^T  with M = 128 output channels
// (A operand = weight tile), N = up to 256 pixels (B operand = plane window): twice the MACs per tcgen05.mma.
// The epilogue thread then owns ONE channel (bias / thresholds are per-thread constants) and its 32-column TMEM
// loads are 32 consecutive pixels, so the 2x2 max pool is register-local.
//
// Replacement table: FMPadding_nonsquare -> TMA OOB zero fill; SWG + stride decimation -> plane offsets; deconv522
// zero insertion -> 4 output phases over one shared plane set; MVAU -> tcgen05.mma kind::i8 (int32 in TMEM);
// PassThrough / bias+ReLU / Thresholds / max pool -> epilogue.
//
// Warps: 0 = weight TMA producer, 1 = MMA issuer (warp-uniform, one elected lane issues), 2..5 = epilogue,
// 6 = plane TMA producer.  plane i: full/empty barrier pair, refilled for the next tile as soon as its last tap has
// retired; weight ring of WSTAGES K-blocks; accumulator stages double buffered in TMEM when they fit.
#include <algorithm>
#include <memory>
#include <vector>

#include "fcb_epilogue.cuh"
#include "fcb_sm100.cuh"
#include "fcb_umma_common.h"

namespace fcb {

using namespace sm100;

// warps: 0 weight TMA, 1 MMA issue, 2..9 epilogue (2 per TMEM lane quarter), 10.. plane TMA producer / im2col builders
constexpr int U2_MAX_KB = 64;
constexpr uint32_t U2_NPB = 4;  // raw patch buffers of the thin-input mode
constexpr int U2_MAX_PLANES = 8;
enum { KB_WAIT = 2, KB_FREE = 4 };

struct KB2 {
  uint16_t plane, flags;
  int32_t a_off;   // plane row offset of this tap's operand (rows of 128 B)
  int32_t w_k;     // K coordinate of the weight tile
  uint32_t d_off;  // (plane smem offset + a_off*128) >> 4 : added to the base smem descriptor
};
struct Plane2 {
  int32_t smem_off, bytes, c0, dx, par, dy, map, pad_;
};
struct Phase2 {
  int32_t nkb, px, py, pad_;
  KB2 kb[U2_MAX_KB];
};
struct Params2 {
  uint8_t* out;
  EpiParams epi;
  int OFM, CB, NPX, stride2, deconv, nphases, nplanes;
  int n_mma;  // N of the resident-planes MMA (<= NPX)
  int WT, R, P, tiles_x, tiles_y, PX, PY;
  int out_x, out_y, out_word_bytes, n_images;
  int wstages, w_bytes, acc_stages, acc_stride, tmem_cols;
  int chb;      // channel blocks the grid is split into (CTA c serves channel block c % chb); > 1 only with smem thresholds
  int thr_off;  // smem offset of the top `thr_top` search levels of this CTA's channels, [2^thr_top - 1][CB*128] (-1: none)
  int thr_top;
  int lut_off;  // smem offset of the bucket LUT of this CTA's channels, [CB*128][256] bytes (-1: none; needs thr_top == all levels)
  int nsets, set_bytes;  // plane sets (double buffering of the input patch across tiles when shared memory allows)
  int w_off, bar_off, stage_off;  // stage_off: 8 x 256 B staging rows of the thin-output epilogue (OFM <= 8)
  // staged epilogue: the tile's output words are assembled in shared memory ([CB][NPX] rows of 128 B, stg_bufs buffers) and
  // leave through TMA stores, one per tile row (whole 128-byte lines, clipped at the image edge by the tensor map)
  int stg_off, stg_bufs, stg_bytes;
  // thin-input mode (input word = one 4-byte pixel of <= 4 lanes, e.g. the C = 3 first layer): warp 10 builds the tile's
  // im2col rows in shared memory (one 128-byte SWIZZLE_128B row per output pixel, word (ky*KX + kx) = input pixel of that tap)
  // from a raw patch fetched by TMA; the layer is then ONE K-block of `ksteps` MMAs per tile with resident weights.
  // swapped orientation (thin-input + bias/ReLU): pixels on the MMA M axis (D[pixel][channel]) so an epilogue thread owns ONE
  // pixel and assembles its 128-byte output word with 16-byte shared-memory stores (the tile is store-bound: K is tiny)
  int swap;
  uint32_t idesc_swap;
  // warp-local stores (swap + alternating groups + WT % 32 == 0): a warp owns whole 128-byte rows of 32 consecutive pixels of one
  // tile row, so it issues its own TMA store per 128-pixel block and the epilogue needs no block-wide barrier at all
  int wl;
  // alternating epilogue groups (staged paths with two accumulator stages): warps 2..5 own accumulator stage / staging buffer 0,
  // warps 6..9 stage 1, so one group's barriers, fences and store issue overlap the other group's TMEM reads and packing
  int epi_alt;
  // MMA-bound layers on the register epilogue: only warps 2..5 work (one per TMEM lane quarter); the second epilogue warp of each
  // SM sub-partition would only compete with the MMA issue warp for issue slots
  int epi4;
  // bias folded into the GEMM (thin-input mode): window word `bias_word` of every im2col row is the constant 1 and the weight
  // byte there is the channel's bias, so the accumulator already is acc + bias (all arithmetic is mod 2^8); -1 = not folded
  int bias_word;
  // thin-output transposed conv (deconv522 with OFM <= 4, e.g. the 3-channel last layer): pixels on the MMA M axis, the 25 taps
  // regrouped by their 9 input shifts so ONE N = 16 instruction serves the 4 output phases x 4 channel slots of a shift:
  // D[pixel][phase*4 + ch]; the epilogue thread owns one input pixel = a 2x2 block of output words
  int dthin;
  uint32_t idesc_dthin;
  // dthin == 2 (col2im form): D[(tap*OFM + ch)][pixel] over the tile's (R+2) x (WT+2) pixels; the epilogue writes the rows as bytes
  // to shared memory (pitch `s_pitch`) and each thread then sums the taps of its input pixel's 2x2 output words
  int s_pitch;
  unsigned long long* prof;  // FCB_U2_PROF: per-CTA clock totals [role 0 builder | 1 mma | 2 epilogue][8 segments]
  int thin_in, S, pad, nw, BWp, BHp, patch_off, patch_bytes, ksteps, wstatic;
  int toff[32];  // patch word offset of window word i: ky*BWp + kx
  int debug;  // FCB_U2_DEBUG bitmask (perf decomposition only): 1 no weight TMA, 2 no plane TMA, 4 no stores, 8 no epilogue
  unsigned long long out_img_bytes;
  uint32_t idesc;
  Plane2 planes[U2_MAX_PLANES];
  Phase2 phases[4];
};

__device__ __forceinline__ bool chv_of(int ch, int ofm) { return ch < ofm; }

// tensor maps of one (input, output, batch) triple: encoding three maps costs ~2 us of host time, and steady-state callers (the
// layer chain's intermediates, the staging slots of the host-buffer calls, benchmarks) present the same few triples again and again
struct MapSet {
  const void* d_in = nullptr;
  void* d_out = nullptr;
  int n_images = 0;
  CUtensorMap tmA[2], tmO;
};
struct Umma2Plan {
  Geom g;
  Params2 p;
  CUtensorMap tmW;
  CUtensorMap tmW_slice;  // box of half the channel rows: the slice one CTA of a 2-CTA cluster fetches and multicasts (cluster_ok)
  bool cluster_ok = false;
  size_t smem;
  int num_sms;
  uint32_t box_rows[2];
  MapSet maps[4];
  int map_next = 0;
};

// pixel bookkeeping of the epilogue: column m of the accumulator -> output word
struct PixMap {
  const Params2& p;
  int x0, y0, px, py;
  unsigned long long img_off;
  __device__ __forceinline__ bool inside(int rr, int xo) const {
    return xo < p.WT && rr < p.R && (x0 + xo) < p.PX && (y0 + rr) < p.PY;
  }
  // byte offset of the output word of tile pixel (rr, xo); pk = pool size
  __device__ __forceinline__ size_t word_off(int rr, int xo, int pk) const {
    const int gx = x0 + xo, gy = y0 + rr;
    const int ox = p.deconv ? 2 * gx + px : gx, oy = p.deconv ? 2 * gy + py : gy;
    return img_off + ((size_t)(oy / pk) * p.out_x + (ox / pk)) * p.out_word_bytes;
  }
  // the same for 2x2-pooled conv layers (no deconv phases), in 32-bit arithmetic below the image offset (an image is < 4 GB)
  __device__ __forceinline__ size_t pooled2_off(int rr, int xo) const {
    return img_off + (unsigned)(((unsigned)((y0 + rr) >> 1) * (unsigned)p.out_x + (unsigned)((x0 + xo) >> 1)) * (unsigned)p.out_word_bytes);
  }
};

// Persistent-tile iterator: CTA c visits tiles c, c + ncta, ... ; (image, tile column, tile row) advance by precomputed
// increments, so the per-tile bookkeeping of every warp role is a handful of adds instead of 32/64-bit divisions.
struct TileIter {
  int img, tx, ty, da, dx, dy, tiles_x, tiles_y;
  long long t, ncta, total;
  __device__ __forceinline__ TileIter(long long cta0, long long ncta_, int tiles_x_, int tiles_y_, int n_images)
      : tiles_x(tiles_x_), tiles_y(tiles_y_), t(cta0), ncta(ncta_), total((long long)tiles_x_ * tiles_y_ * n_images) {
    const uint32_t tpi = (uint32_t)(tiles_x * tiles_y), c = (uint32_t)cta0, n = (uint32_t)ncta_;
    img = (int)(c / tpi);
    const uint32_t r = c - (uint32_t)img * tpi;
    ty = (int)(r / (uint32_t)tiles_x); tx = (int)(r - (uint32_t)ty * tiles_x);
    da = (int)(n / tpi);
    const uint32_t b = n - (uint32_t)da * tpi;
    dy = (int)(b / (uint32_t)tiles_x); dx = (int)(b - (uint32_t)dy * tiles_x);
  }
  __device__ __forceinline__ bool valid() const { return t < total; }
  __device__ __forceinline__ void next() {
    t += ncta; tx += dx; ty += dy; img += da;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++img; }
  }
};
// ring of n plane sets walked tile by tile (n is small but not a power of two in general: no division per tile)
struct SetRing {
  int idx = 0, n;
  uint32_t par = 0;
  __device__ __forceinline__ explicit SetRing(int n_) : n(n_) {}
  __device__ __forceinline__ void next() { if (++idx == n) { idx = 0; par ^= 1u; } }
};
// rings of 1 or 2 slots: slot and phase parity of iteration `it`
__device__ __forceinline__ uint32_t ring_idx(uint32_t it, int n) { return n == 2 ? (it & 1u) : 0u; }
__device__ __forceinline__ uint32_t ring_par(uint32_t it, int n) { return n == 2 ? ((it >> 1) & 1u) : (it & 1u); }

// One im2col row of the thin-input mode: 8*KS window words gathered from the raw patch (offsets in registers), written as 2*KS
// 16-byte chunks of a SWIZZLE_128B row.  Straight-line: all loads are issued before the first store.
template <int KS, int BW>
__device__ __forceinline__ void build_row(uint32_t src, uint32_t dst, uint32_t m7, const uint32_t (&offs)[32]) {
  uint32_t w[8 * KS];
#pragma unroll
  for (int i = 0; i < 8 * KS; i++) {
    if (i == BW) w[i] = 1u;  // the constant-1 activation that multiplies the bias row of the weights
    else if (BW >= 0 && i > BW) w[i] = 0u;
    else w[i] = lds_b32(src + offs[i]);
  }
#pragma unroll
  for (int j = 0; j < 2 * KS; j++) sts_v4(dst + (((uint32_t)j ^ m7) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

// Two horizontally adjacent im2col rows of a 5x5 stride-2 window (the first layer of the reference net) at once: pixel x+1's
// window starts two patch words after pixel x's, so the two rows share 3 of the 5 words of every window line.  Per window line
// the thread reads 8 consecutive patch words as four 8-byte loads (src is 8-byte aligned: even pad, even x) instead of 2 x 5
// single words: 20 loads for two rows instead of 50.  Row layout as build_row<4, 25>: 25 window words, the constant 1 of the bias
// row, zeros.
__device__ __forceinline__ void build_row_pair_5x5s2(uint32_t src, uint32_t line_step, uint32_t dst, uint32_t m7) {
  uint32_t a[32], b[32];
#pragma unroll
  for (int ky = 0; ky < 5; ky++) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint2 v = lds_v2(src + (uint32_t)ky * line_step + 8u * (uint32_t)i);
      w[2 * i] = v.x; w[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int kx = 0; kx < 5; kx++) { a[ky * 5 + kx] = w[kx]; b[ky * 5 + kx] = w[kx + 2]; }
  }
  a[25] = b[25] = 1u;
#pragma unroll
  for (int i = 26; i < 32; i++) a[i] = b[i] = 0u;
  const uint32_t m7b = m7 + 1u;  // (m is even: the second row's swizzle phase is the next one, no wrap inside the pair)
#pragma unroll
  for (int j = 0; j < 8; j++) sts_v4(dst + (((uint32_t)j ^ m7) << 4), a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
#pragma unroll
  for (int j = 0; j < 8; j++) sts_v4(dst + 128u + (((uint32_t)j ^ m7b) << 4), b[4 * j], b[4 * j + 1], b[4 * j + 2], b[4 * j + 3]);
}

// col2im of the thin-output transposed conv.  The byte tile is PIXEL-major: S[pixel][word][ch], one 4-byte word per tap
// (word order dcol_word(), fcb_internal.h: taps grouped by input shift, sorted by output phase inside a group), pitch DCOL_PIX bytes
// per pixel.  An input pixel's 2x2 output block is then 9 vector loads (one per shift: 4 x 16 B, 4 x 8 B, 1 x 4 B) and 21 packed
// byte adds instead of 75 byte loads and adds.  pitch 112 B = 28 words: the 16-byte loads of 8 consecutive pixels hit 8 bank groups.
constexpr int DCOL_PIX = 112;
template <int IMM>
__device__ __forceinline__ void lds_v4_imm(uint32_t base, uint32_t (&v)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4 + %5];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(base), "n"(IMM));
}
template <int IMM>
__device__ __forceinline__ void lds_v2_imm(uint32_t base, uint32_t (&v)[2]) {
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2 + %3];" : "=r"(v[0]), "=r"(v[1]) : "r"(base), "n"(IMM));
}
template <int IMM>
__device__ __forceinline__ uint32_t lds_b32_imm(uint32_t base) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1 + %2];" : "=r"(v) : "r"(base), "n"(IMM));
  return v;
}
// bm[i] = shared address of pixel (r + i - 1, x - 1): shift (oy, ox) is bm[oy + 1] + (ox + 1) * DCOL_PIX.  s[ph] = packed byte
// sums of output phase ph = 2 * py + px (all arithmetic mod 2^8 per byte lane).
__device__ __forceinline__ void dcol_gather(const uint32_t (&bm)[3], uint32_t (&s)[4]) {
  uint32_t g[4], h[2];
  lds_v4_imm<1 * DCOL_PIX + 0>(bm[1], s);                      // shift ( 0,  0): ky in {1,2}, kx in {1,2}
  lds_v4_imm<2 * DCOL_PIX + 16>(bm[1], g);                     // shift ( 0, +1)
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = __vadd4(s[i], g[i]);
  lds_v4_imm<1 * DCOL_PIX + 32>(bm[2], g);                     // shift (+1,  0)
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = __vadd4(s[i], g[i]);
  lds_v4_imm<2 * DCOL_PIX + 48>(bm[2], g);                     // shift (+1, +1)
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = __vadd4(s[i], g[i]);
  lds_v2_imm<1 * DCOL_PIX + 64>(bm[0], h);                     // shift (-1,  0): ky = 0 -> py = 0; words = px 0, 1
  s[0] = __vadd4(s[0], h[0]); s[1] = __vadd4(s[1], h[1]);
  lds_v2_imm<2 * DCOL_PIX + 72>(bm[0], h);                     // shift (-1, +1)
  s[0] = __vadd4(s[0], h[0]); s[1] = __vadd4(s[1], h[1]);
  lds_v2_imm<0 * DCOL_PIX + 80>(bm[1], h);                     // shift ( 0, -1): kx = 0 -> px = 0; words = py 0, 1
  s[0] = __vadd4(s[0], h[0]); s[2] = __vadd4(s[2], h[1]);
  lds_v2_imm<0 * DCOL_PIX + 88>(bm[2], h);                     // shift (+1, -1)
  s[0] = __vadd4(s[0], h[0]); s[2] = __vadd4(s[2], h[1]);
  s[0] = __vadd4(s[0], lds_b32_imm<0 * DCOL_PIX + 96>(bm[0])); // shift (-1, -1): tap (0, 0)
}

// NB = warps from index 10 on: 1 = plane TMA producer (resident-planes mode); 4 = im2col builders (thin-input mode)
// DTHIN: the thin-output transposed-conv instantiation.  Mode flags are template parameters because the MMA issue loop has no
// slack for per-K-block loads of kernel parameters (each LDCU + dependent branch costs ~50 clocks of its ~580-clock budget).
// EPI = 1: the instantiation for pooled 8-bit threshold layers (monotone compare, shared-memory tables): every other epilogue is
// compiled out, which frees registers and instruction cache for the lock-step search.
// CS = CTAs per cluster (1 or 2) sharing ONE stream of weight K-blocks: each CTA fetches 1/CS of every K-block and multicasts it into
// every ring of the cluster (the L2 -> SM weight stream -- 400 KB per 256-pixel tile on CONV_1 -- was the largest data mover of the
// channel-heavy layers and, under the 1 kW power cap, clocks: profiles/r02_power_decomposition.log).  Each CTA still runs its own
// tiles with cta_group::1 MMAs; only the weight ring is shared, so all CTAs of a cluster step through the same number of K-blocks
// (a CTA with one tile fewer runs the ring protocol of a phantom tile).
template <int NB, int DT, int EPI, int CS = 1>
__global__ void __launch_bounds__(320 + 32 * NB + (EPI >= 4 ? 256 : EPI >= 2 ? 128 * (EPI - 1) : 0), 1)
umma2_conv_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const Params2 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* wfull = bars;
  uint64_t* wempty = wfull + p.wstages;
  uint64_t* afull = wempty + p.wstages;
  uint64_t* aempty = afull + p.nsets * p.nplanes;
  uint64_t* tfull = aempty + p.nsets * p.nplanes;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* pfull = tempty + 3;  // thin-input mode: raw patch buffers (U2_NPB)
  uint64_t* pempty = pfull + U2_NPB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool THIN = NB > 1;         // thin-input instantiations (im2col builder warps)
  constexpr bool THRP = EPI >= 1 && EPI <= 3;   // pooled 8-bit thresholds only
  // EPI = 4: pixel-major bias+ReLU with warp-local TMA stores only (thin-input layers), FOUR epilogue groups: group g serves
  // accumulator stage g & 1 and the 128-pixel block g >> 1 of its tiles (a warp issues one instruction per ~5 clocks on this
  // path -- fixed-latency dependency stalls -- so the issue slots are filled with more warps instead)
  constexpr bool SWPX = EPI == 4;
  constexpr bool GEN = EPI == 0;  // the general instantiation: every epilogue family behind run-time flags
  // EPI = 5 (with DT = 2): the col2im epilogue alone, four groups: groups g and g + 2 share accumulator stage g & 1 and its byte
  // tile, splitting the columns of phase 1 and the pixels of phase 2
  constexpr bool DCX = EPI == 5;
  // EPI = 6: the staged channel-major bias+ReLU epilogue alone (two staging tiles, one per accumulator stage), four groups:
  // groups g and g + 2 share stage g & 1, taking alternate 32-column blocks and alternate rows of the TMA stores
  constexpr bool STX = EPI == 6;
  constexpr int XEPI = (SWPX || DCX || STX) ? 8 : EPI >= 2 ? 4 * (EPI - 1) : 0;  // EPI = 2, 3: 4 / 8 more epilogue warps behind the producer warps (the search is latency-bound)
  constexpr bool DTHIN = DT == 1;   // thin-output deconv, 9 shift blocks x N=16 (pixels on M)
  constexpr bool DCOL = DT == 2;    // thin-output deconv, GEMM over (tap, channel) rows + col2im in shared memory
  constexpr bool WSTATIC = THIN || DT != 0;  // every weight K-block has its own stage: loaded once, never released
#define WAITB(BAR_, PAR_) mbar_wait((BAR_), (PAR_))
  // clock accounting and the perf-decomposition switches of the inner loops exist only in -DFCB_U2_PROF builds (tools/): even a
  // predicted-not-taken branch per K-block shows in the MMA issue loop, which has ~580 clocks per K-block to stay ahead of the pipe
#if defined(FCB_U2_PROF)
  long long prof_c = 0;
  unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_START() do { if (p.prof) prof_c = clock64(); } while (0)
#define PROF_T(SEG_) do { if (p.prof) { const long long n_ = clock64(); prof_acc[SEG_] += (unsigned long long)(n_ - prof_c); prof_c = n_; } } while (0)
#define PROF_FLUSH(ROLE_) do { if (p.prof) for (int i_ = 0; i_ < 8; i_++) p.prof[(blockIdx.x * 3 + (ROLE_)) * 8 + i_] = prof_acc[i_]; } while (0)
#define DBG(MASK_) (p.debug & (MASK_))
#else
#define PROF_START() do { } while (0)
#define PROF_T(SEG_) do { } while (0)
#define PROF_FLUSH(ROLE_) do { } while (0)
#define DBG(MASK_) (false)
#endif
  const long long tiles_per_img = (long long)p.tiles_x * p.tiles_y;
  const long long total_tiles = tiles_per_img * p.n_images;
  // CTA -> (channel block, tile sequence): with chb > 1 every tile is visited once per channel block
  const int chblk = blockIdx.x % p.chb, chbase = chblk * 128;
  const long long cta0 = blockIdx.x / p.chb, ncta = gridDim.x / p.chb;
  // weight-ring passes of this CTA: its own tiles, or (clusters) the tile count of the busiest CTA -- uniform over the cluster
  const long long my_tiles = cta0 < total_tiles ? (total_tiles - cta0 + ncta - 1) / ncta : 0;
  const long long ring_tiles = CS > 1 ? (total_tiles + ncta - 1) / ncta : my_tiles;
  const uint32_t cta_rank = CS > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CS) - 1u);
  if (p.thr_off >= 0) {
    // top levels of the threshold search for this CTA's channels, threshold-major (see activate_thr_hybrid)
    int32_t* ts = reinterpret_cast<int32_t*>(smem + p.thr_off);
    const int nch = p.CB * 128, ntop = (1 << p.thr_top) - 1;
    int gshift = 0;
    for (int t = p.epi.thr_n + 1; (t >> (p.thr_top + gshift)) > 1;) gshift++;
    // pooled-threshold instantiations with a bucket LUT search with ONE strict compare (thr_lut_fast): thr <= a  <=>  thr - 1 < a
    const int32_t thr_adj = (EPI >= 1 && EPI <= 3 && p.lut_off >= 0 && p.epi.cmp == FCB_CMP_LESS_EQUAL) ? 1 : 0;
    for (int idx = threadIdx.x; idx < ntop * nch; idx += blockDim.x) {
      const int j = idx / nch + 1, c = idx - (j - 1) * nch;
      ts[idx] = __ldg(p.epi.thr + (size_t)((j << gshift) - 1) * p.epi.thr_stride + chbase + c) - thr_adj;
    }
  }

  if (p.lut_off >= 0) {
    const uint4* src = reinterpret_cast<const uint4*>(p.epi.thr_lut + (size_t)chbase * 256);
    uint4* dst = reinterpret_cast<uint4*>(smem + p.lut_off);
    for (int idx = threadIdx.x; idx < p.CB * 128 * 16; idx += blockDim.x) dst[idx] = __ldg(src + idx);
  }
  if (p.swap)  // bias bytes of this CTA's channels (read as packed words by every epilogue thread)
    for (int idx = threadIdx.x; idx < p.CB * 128; idx += blockDim.x) smem[p.stage_off + idx] = idx < p.OFM ? (uint8_t)p.epi.bias[idx] : 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.wstages; s++) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], CS); }  // a stage is free when EVERY CTA of the cluster has read it
    for (int i = 0; i < p.nsets * p.nplanes; i++) { mbar_init(&afull[i], p.thin_in ? NB : 1); mbar_init(&aempty[i], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], THRP ? 8 + XEPI : (SWPX || DCX || STX) ? 8 : (p.epi_alt || p.epi4) ? 4 : 8); }
    for (int a = 0; a < U2_NPB; a++) { mbar_init(&pfull[a], 1); mbar_init(&pempty[a], NB); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // peers' barriers are initialised before any remote arrival or multicast write
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: weight K-blocks =====================
    if (lane == 0) {
      int s = 0;
      uint32_t wphase = 1;
      for (long long t = 0; t < ring_tiles; t++) {
        if (WSTATIC && t != 0) break;
        for (int ph = 0; ph < p.nphases; ph++) {
          const Phase2& P = p.phases[ph];
          for (int i = 0; i < P.nkb; i++) {
            mbar_wait(&wempty[s], wphase);
            if (DBG(1)) mbar_arrive(&wfull[s]);
            else {
              mbar_arrive_expect_tx(&wfull[s], (uint32_t)p.w_bytes);  // the whole K-block: 1/CS of it from every CTA of the cluster
              if (DTHIN || DCOL) tma_load_2d(smem + p.w_off + s * p.w_bytes, &tmW, &wfull[s], 0, P.kb[i].w_k);  // row block w_k
              else if (CS > 1)  // this CTA's slice of the channel rows (tmW's box is CB*128/CS rows), into every ring of the cluster
                tma_load_2d_mc(smem + p.w_off + s * p.w_bytes + cta_rank * (uint32_t)(p.w_bytes / CS), &tmW, &wfull[s], P.kb[i].w_k,
                               chbase + (int)cta_rank * (p.CB * 128 / CS), MC_MASK);
              else tma_load_2d(smem + p.w_off + s * p.w_bytes, &tmW, &wfull[s], P.kb[i].w_k, chbase);
            }
            if (++s == p.wstages) { s = 0; wphase ^= 1; }
          }
        }
      }
      if (p.thin_in) {
        // ===== raw patch fetcher (thin-input mode; the weights above were loaded once): U2_NPB buffers, released by the builders
        uint32_t it = 0;
        for (TileIter ti(cta0, ncta, p.tiles_x, p.tiles_y, p.n_images); ti.valid(); ti.next(), it++) {
          const uint32_t b = it % U2_NPB;
          if (it >= U2_NPB) WAITB(&pempty[b], ((it / U2_NPB) & 1u) ^ 1u);
          if (DBG(32)) { mbar_arrive(&pfull[b]); continue; }
          mbar_arrive_expect_tx(&pfull[b], (uint32_t)(p.BWp * p.BHp * 4));
          // the box starts at a multiple of 4 pixels: an un-swizzled TMA box must start 16-byte aligned in its innermost dimension
          // (anything else is an illegal instruction: tools/tma_probe.cu, profiles/r01_tma_inner_alignment_probe.log)
          tma_load_3d(smem + p.patch_off + b * p.patch_bytes, &tmA0, &pfull[b], (p.S * ti.tx * p.WT - p.pad) & ~3,
                      p.S * ti.ty * p.R - p.pad, ti.img);
        }
      }
    }
  } else if (warp >= 10 && warp < 10 + NB) {
    // ===================== TMA producer: input planes =====================
    // Independent of the weight ring: plane i of the next tile is fetched the moment the MMAs that read plane i of
    // the current tile have retired (aempty), i.e. while the remaining taps of the current tile execute.
    if (p.thin_in) {
      // ===== sliding window as shared-memory im2col (slidingwindow.h:1302-1313 order ky, kx, lane; FMPadding zeros = TMA OOB fill)
      // builder warp bw takes rows m = 32*bw + lane (mod 32*NB).  Window word i of a row comes from patch word
      // offs[i] = ky*BWp + kx (bytes, kept in registers); words beyond the window re-read word 0: their weights are zero.
      const int bw = warp - 10;
      uint32_t offs[32];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        offs[i] = i < p.nw ? 4u * (uint32_t)p.toff[i] : 0u;
        asm volatile("" : "+r"(offs[i]));
      }
      const int npix = p.R * p.WT, ksteps = p.ksteps, bias_word = p.bias_word;
      // the 5x5 stride-2 window with the bias row (L0): rows are built in pairs; needs 8-byte aligned window starts (even pad, even
      // tile width) and an even number of pixels per tile line
      const bool pair55 = bias_word == 25 && p.S == 2 && p.nw == 25 && p.toff[4] == 4 && p.toff[5] == p.BWp && !(p.pad & 1) && !(p.WT & 1) && !DBG(256);
      const uint32_t row_step = 4u * (uint32_t)(p.S * p.BWp), col_step = 4u * (uint32_t)p.S;
      uint32_t tile_it = 0;
      PROF_START();
      SetRing sr(p.nsets);
      for (TileIter ti(cta0, ncta, p.tiles_x, p.tiles_y, p.n_images); ti.valid(); ti.next(), tile_it++, sr.next()) {
        const uint32_t set = (uint32_t)sr.idx, pb = tile_it % U2_NPB;
        PROF_T(0);
        WAITB(&pfull[pb], (tile_it / U2_NPB) & 1u);
        PROF_T(1);
        WAITB(&aempty[set * p.nplanes], sr.par ^ 1u);
        PROF_T(2);
        const uint32_t xshift = (uint32_t)((p.S * ti.tx * p.WT - p.pad) & 3);
        const uint32_t patch = smem_u32(smem + p.patch_off + pb * p.patch_bytes) + 4u * xshift;
        const uint32_t rows = smem_u32(smem + set * p.set_bytes + p.planes[0].smem_off);
        if (pair55) {
          // rows 2j, 2j+1 (same tile line: WT is even), j = 32*bw + lane (+ 32*NB per pass)
          int m = 2 * (32 * bw + lane);
          int rr = m / p.WT, xo = m - rr * p.WT;
          for (; m < (DBG(64) ? 0 : npix); m += 64 * NB) {
            const uint32_t src = patch + (uint32_t)rr * row_step + (uint32_t)xo * col_step;
            build_row_pair_5x5s2(src, 4u * (uint32_t)p.BWp, rows + 128u * (uint32_t)m, (uint32_t)m & 7u);
            xo += 64 * NB;
            while (xo >= p.WT) { xo -= p.WT; ++rr; }
          }
        } else {
        int rr = (32 * bw + lane) / p.WT, xo = (32 * bw + lane) - rr * p.WT;
        for (int m = 32 * bw + lane; m < (DBG(64) ? 0 : npix); m += 32 * NB) {
          uint32_t src = patch + (uint32_t)rr * row_step + (uint32_t)xo * col_step;
          asm volatile("" : "+r"(src));  // one register, not re-derived per load
          const uint32_t dst = rows + 128u * (uint32_t)m, m7 = (uint32_t)m & 7u;
          if (bias_word == 25) build_row<4, 25>(src, dst, m7, offs);     // 5x5 window
          else if (bias_word == 9) build_row<2, 9>(src, dst, m7, offs);  // 3x3 window
          else if (ksteps == 4) build_row<4, -1>(src, dst, m7, offs);
          else if (ksteps == 2) build_row<2, -1>(src, dst, m7, offs);
          else if (ksteps == 3) build_row<3, -1>(src, dst, m7, offs);
          else build_row<1, -1>(src, dst, m7, offs);
          xo += 32 * NB;
          while (xo >= p.WT) { xo -= p.WT; ++rr; }
        }
        }
        PROF_T(3);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&afull[set * p.nplanes]); mbar_arrive(&pempty[pb]); }
        PROF_T(4);
      }
      if (bw == 0 && lane == 0) PROF_FLUSH(0);
    } else if (lane == 0) {
      uint32_t tile_it = 0;
      SetRing sr(p.nsets);
      for (TileIter ti(cta0, ncta, p.tiles_x, p.tiles_y, p.n_images); ti.valid(); ti.next(), tile_it++, sr.next()) {
        const int img = ti.img, x0 = ti.tx * p.WT, y0 = ti.ty * p.R;
        const int set = sr.idx;
        const uint32_t par = sr.par ^ 1u;
        for (int i = 0; i < p.nplanes; i++) {
          const Plane2& pl = p.planes[i];
          uint64_t* fb = &afull[set * p.nplanes + i];
          mbar_wait(&aempty[set * p.nplanes + i], par);
          if (DBG(2)) { mbar_arrive(fb); continue; }
          mbar_arrive_expect_tx(fb, (uint32_t)pl.bytes);
          const CUtensorMap* m = pl.map ? &tmA1 : &tmA0;
          uint8_t* dst = smem + set * p.set_bytes + pl.smem_off;
          if (p.stride2) tma_load_5d(dst, m, fb, pl.c0, x0 + pl.dx, pl.par, y0 + pl.dy, img);
          else tma_load_4d(dst, m, fb, pl.c0, x0 + pl.dx, y0 + pl.dy, img);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop in uniform control flow (addresses and descriptors stay in uniform registers,
    // which UTCIMMA consumes); one elected lane issues the tcgen05 instructions.
    // Per K-block work is kept to: poll the weight stage, add two precomputed offsets to the base descriptor, issue.
    uint32_t tile_it = 0, acc_it = 0;
    int s = 0;
    uint32_t wphase = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = p.idesc;
    const uint64_t desc0 = make_smem_desc(smem_u32(smem), 128);
    const uint32_t w_d0 = (uint32_t)p.w_off >> 4, w_dstep = (uint32_t)p.w_bytes >> 4;
    int wstages = p.wstages, CB = p.CB;
    asm volatile("" : "+r"(wstages), "+r"(CB));  // held in registers: not re-read from the parameter bank per K-block
    const int NPX = p.NPX, ksteps = p.ksteps;
    PROF_START();
    SetRing sr(p.nsets);
    for (long long t = cta0; t < total_tiles; t += ncta, tile_it++, sr.next()) {
      const int set = sr.idx;
      const uint32_t apar = sr.par;
      const uint64_t desc_set = desc0 + (uint32_t)((set * p.set_bytes) >> 4);
      for (int ph = 0; ph < p.nphases; ph++, acc_it++) {
        const Phase2& P = p.phases[ph];
        const int nkb = P.nkb;
        const int acc = (int)ring_idx(acc_it, p.acc_stages);
        PROF_T(0);
        WAITB(&tempty[acc], ring_par(acc_it, p.acc_stages) ^ 1u);
        PROF_T(1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * p.acc_stride);
        for (int i = 0; i < nkb; i++) {
          const uint32_t flags = P.kb[i].flags, plane = set * p.nplanes + P.kb[i].plane;
          const uint64_t pdesc = desc_set + P.kb[i].d_off;
          PROF_T(2);
          if (flags & KB_WAIT) WAITB(&afull[plane], apar);
          PROF_T(3);
          if (!WSTATIC || tile_it == 0) mbar_wait(&wfull[s], wphase);  // resident weights: loaded once, observed once
          PROF_T(4);
          tc_fence_after();
          const uint64_t wdesc = desc0 + (w_d0 + (uint32_t)s * w_dstep);
          if (elect_one_sync()) {
            if (DTHIN || DCOL) {
              // A = 128 plane rows (pixels) per block at this shift, B = 16 weight rows (4 phases x 4 channel slots);
              // col2im form: B = the 112 (tap word, channel) weight rows, D[pixel][tap word * 4 + channel]
              constexpr int DSTEP = DCOL ? 112 : 16;
              for (int blk = 0; blk < NPX / 128; blk++) {
                const uint32_t dt = d_tmem + (uint32_t)(blk * DSTEP);
                const uint64_t ad = pdesc + (uint64_t)(blk * 1024);
                umma_i8(dt, ad, wdesc, p.idesc_dthin, i ? 1u : 0u);
                umma_i8(dt, ad + 2, wdesc + 2, p.idesc_dthin, 1u);
                umma_i8(dt, ad + 4, wdesc + 4, p.idesc_dthin, 1u);
                umma_i8(dt, ad + 6, wdesc + 6, p.idesc_dthin, 1u);
              }
            } else if (THIN && p.swap) {
              // A = 128 im2col rows (pixels) per block, B = the CB*128 weight rows: D[pixel][channel]
              if (!DBG(128))
                for (int blk = 0; blk < NPX / 128; blk++) {
                  const uint32_t dt = d_tmem + (uint32_t)(blk * CB * 128);
                  const uint64_t ad = pdesc + (uint64_t)(blk * 1024);  // 128 rows x 128 B = 16 KB (>> 4)
                  umma_i8(dt, ad, wdesc, p.idesc_swap, 0u);
                  if (ksteps > 1) umma_i8(dt, ad + 2, wdesc + 2, p.idesc_swap, 1u);
                  if (ksteps > 2) umma_i8(dt, ad + 4, wdesc + 4, p.idesc_swap, 1u);
                  if (ksteps > 3) umma_i8(dt, ad + 6, wdesc + 6, p.idesc_swap, 1u);
                }
            } else if (!DBG(128)) {
              umma_i8(d_tmem, wdesc, pdesc, idesc, i ? 1u : 0u);
              if (!THIN || ksteps > 1) umma_i8(d_tmem, wdesc + 2, pdesc + 2, idesc, 1u);
              if (!THIN || ksteps > 2) umma_i8(d_tmem, wdesc + 4, pdesc + 4, idesc, 1u);
              if (!THIN || ksteps > 3) umma_i8(d_tmem, wdesc + 6, pdesc + 6, idesc, 1u);
              if (CB == 2) {
                const uint32_t dt = d_tmem + (uint32_t)NPX;
                umma_i8(dt, wdesc + 1024, pdesc, idesc, i ? 1u : 0u);  // next 128 weight rows = +16 KB (>> 4)
                if (!THIN || ksteps > 1) umma_i8(dt, wdesc + 1026, pdesc + 2, idesc, 1u);
                if (!THIN || ksteps > 2) umma_i8(dt, wdesc + 1028, pdesc + 4, idesc, 1u);
                if (!THIN || ksteps > 3) umma_i8(dt, wdesc + 1030, pdesc + 6, idesc, 1u);
              }
            }
            if (!WSTATIC) { if (CS > 1) umma_commit_mc(&wempty[s], MC_MASK); else umma_commit(&wempty[s]); }
            if (flags & KB_FREE) umma_commit(&aempty[plane]);
            if (i == nkb - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++s == wstages) { s = 0; wphase ^= WSTATIC ? 0u : 1u; }
        }
      }
    }
    if (CS > 1) {
      // phantom passes: a CTA with fewer tiles than the busiest one of its cluster still takes every K-block off the shared ring
      // (its stages are written by the peers' multicasts) and frees it, without issuing MMAs
      for (long long t = my_tiles; t < ring_tiles; t++)
        for (int ph = 0; ph < p.nphases; ph++)
          for (int i = 0; i < p.phases[ph].nkb; i++) {
            mbar_wait(&wfull[s], wphase);
            if (elect_one_sync()) umma_commit_mc(&wempty[s], MC_MASK);
            __syncwarp();
            if (++s == wstages) { s = 0; wphase ^= 1u; }
          }
    }
    if (lane == 0) PROF_FLUSH(1);
  } else if (warp < 10 || XEPI > 0) {
    // ===================== epilogue (warps 2..9): lane = output channel, TMEM column = pixel =====================
    // two warps per TMEM lane quarter; `half` splits the accumulator columns (pixels) between them
    // (EPI >= 2: further groups of four warps behind the producers, `half` = 2, 3)
    const int q = warp & 3, half = warp < 10 ? (warp - 2) >> 2 : 2 + ((warp - 10 - NB) >> 2);
    constexpr int NHALF = 2 + XEPI / 4;
    const bool alt = SWPX || DCX || STX || (GEN && p.epi_alt != 0);
    const bool whole = alt || (GEN && p.epi4);  // this warp covers every column of the accumulators it serves
    const int col_lo = STX ? 32 * (half >> 1) : whole ? 0 : half * (p.NPX / 2), col_hi = whole ? p.NPX : col_lo + p.NPX / 2;
    constexpr int col_step = STX ? 64 : 32;
    const int ebar = alt ? 1 + (half & 1) : 1, ecnt = (DCX || STX) ? 256 : alt ? 128 : 256;      // named barrier of this warp's epilogue group
    const int erow0 = STX ? 4 * (half >> 1) + q : alt ? q : warp - 2, erows = (alt && !STX) ? 4 : 8;        // TMA-store rows dealt over the group's warps
#define EPI_BAR() asm volatile("bar.sync %0, %1;" ::"r"(ebar), "r"(ecnt) : "memory")
    const int pk = THRP ? 2 : (p.epi.pool >= 2 ? p.epi.pool : 1);
    const bool fast = GEN && p.epi.act_kind == FCB_ACT_BIAS_RELU && p.epi.out_bits == 8 && p.epi.acc_bits == 8 && pk == 1 && (p.OFM % 32) == 0 && p.P >= 4;
    // thresholds with comp::less / less_equal and a result that cannot wrap TR: monotone in the (wrapped, < 2^31) accumulator
    const bool mono = THRP || p.epi.act_kind == FCB_ACT_THRESHOLDS && (p.epi.cmp == FCB_CMP_LESS || p.epi.cmp == FCB_CMP_LESS_EQUAL) &&
                      p.epi.act_val >= 0 && (p.epi.out_bits >= 31 || p.epi.act_val + p.epi.num_th < (1 << p.epi.out_bits)) &&
                      (p.epi.acc_signed || p.epi.acc_bits < 32);
    int gshift = 0;  // log2 of the group the shared-memory levels of the threshold search narrow down to
    if (p.thr_off >= 0) for (int t = p.epi.thr_n + 1; (t >> (p.thr_top + gshift)) > 1;) gshift++;
    const bool thin = GEN && p.epi.act_kind == FCB_ACT_BIAS_RELU && p.epi.out_bits == 8 && p.epi.acc_bits == 8 && pk == 1 && p.OFM <= 8;
    uint32_t acc_it = 0;
    uint32_t dcol_bias = 0;  // col2im epilogue: the output word's bias bytes
    if (DCOL)
      for (int o = 0; o < p.OFM && o < 4; o++) dcol_bias |= ((uint32_t)(int32_t)p.epi.bias[o] & 0xFFu) << (8 * o);
    PROF_START();
    // alternating groups on single-phase layers own every other tile outright: they walk their own tile sequence (stride 2 CTAs'
    // worth) instead of stepping through -- and skipping -- the other group's tiles (~40 instructions at 6-8 clocks each per tile)
    // tile-invariant pixel coordinates of this warp's TMA-store block (EPI = 4) / this thread's col2im pixel: divided once here
    int pre_rr = 0, pre_xo = 0;
    if (SWPX || DCOL) {
      const int i0 = SWPX ? (half >> 1) * 128 + q * 32 : (DCX ? (half >> 1) * 128 : 0) + (warp & 3) * 32 + lane;
      pre_rr = i0 / p.WT;
      pre_xo = i0 - pre_rr * p.WT;
    }
    // bucket-LUT constants of this thread's channel(s): tile-invariant, fetched once (two global loads per tile and warp otherwise)
    int32_t lutlo0 = 0, lutlo1 = 0;  // (scalars, not an indexed array: no local-memory frame in the epilogue)
    int lutsh0 = 0, lutsh1 = 0;
    if (p.lut_off >= 0) {
      const int cha = chbase + q * 32 + lane, chb_ = cha + 128;
      lutlo0 = p.epi.thr_lo[chv_of(cha, p.OFM) ? cha : 0];
      lutsh0 = p.epi.thr_sh[chv_of(cha, p.OFM) ? cha : 0];
      if (p.CB > 1) {
        lutlo1 = p.epi.thr_lo[chv_of(chb_, p.OFM) ? chb_ : 0];
        lutsh1 = p.epi.thr_sh[chv_of(chb_, p.OFM) ? chb_ : 0];
      }
    }
    const bool alt1 = alt && p.nphases == 1;
    const uint32_t acc_step = alt1 ? 2u : 1u;
    if (alt1) acc_it = (uint32_t)(half & 1);
    for (TileIter ti(alt1 ? cta0 + (half & 1) * ncta : cta0, alt1 ? 2 * ncta : ncta, p.tiles_x, p.tiles_y, p.n_images);
         ti.valid() && !(GEN && p.epi4 && half); ti.next()) {
      const int img = ti.img;
      for (int ph = 0; ph < p.nphases; ph++, acc_it += acc_step) {
        const int acc = (int)ring_idx(acc_it, p.acc_stages);
        if (alt && acc != (half & 1)) continue;  // the other group's accumulator
        const PixMap pm{p, ti.tx * p.WT, ti.ty * p.R, p.phases[ph].px, p.phases[ph].py, (unsigned long long)img * p.out_img_bytes};
        PROF_T(0);
        WAITB(&tfull[acc], ring_par(acc_it, p.acc_stages));
        PROF_T(1);
        tc_fence_after();
        // valid extent of this tile (rows/columns past it are halo, padding or beyond the image)
        const int vrows = DBG(4) ? 0 : min(p.R, p.PY - pm.y0), vcols = min(p.WT, p.PX - pm.x0);
        if ((GEN || DCX) && DCOL) {
          // (always two accumulator stages: warps 2..5 serve stage 0, warps 6..9 stage 1, each group with its own byte tile)
          // (25 taps x 4 byte lanes = 100 accumulator columns; lanes >= OFM have zero weights: they give the zero pad byte)
          uint8_t* S = smem + p.stg_off + (half & 1) * p.stg_bytes;
          const uint32_t S_s = smem_u32(S);
          const int sub = DCX ? half >> 1 : 0;  // which of the stage's two groups this warp belongs to
          constexpr int NSUB = DCX ? 2 : 1;
          EPI_BAR();  // the group's previous col2im pass has finished reading S
          PROF_T(2);
          // ---- phase 1: accumulator rows -> bytes (all arithmetic is mod 2^8, so the partial sums may be truncated now)
          if (!DBG(8)) {
            // thread = pixel (TMEM lane), columns = the bytes of its record: 3 x 32 columns -> 2 x 16-byte stores each, + the
            // last word (tap (0,0)); the stage's groups (two with 16 epilogue warps) take one 128-pixel block each
            const int ncols = (vrows + 2) * p.P;  // pixels of the tile, halo ring included
            for (int blk = sub; blk < 2; blk += NSUB) {
              if (blk * 128 + q * 32 >= ncols) continue;  // warp-uniform
              const int m = blk * 128 + q * 32 + lane;
              const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride + blk * 112);
              const uint32_t srec = S_s + (uint32_t)(m * DCOL_PIX);
#pragma unroll 1
              for (int cbk = 0; cbk < 3; cbk++) {
                uint32_t v[32];
                tmem_ld32(tacc + (uint32_t)(cbk * 32), v);
                tmem_ld_wait();
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  const uint32_t lo = __byte_perm(v[4 * j], v[4 * j + 1], 0x4040), hi = __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x4040);
                  w[j] = __byte_perm(lo, hi, 0x5410);  // low bytes = accumulators mod 2^8
                }
                if (m < ncols) {
                  sts_v4(srec + (uint32_t)(cbk * 32), w[0], w[1], w[2], w[3]);
                  sts_v4(srec + (uint32_t)(cbk * 32) + 16u, w[4], w[5], w[6], w[7]);
                }
              }
              uint32_t t[8];
              tmem_ld8(tacc + 96u, t);
              tmem_ld_wait();
              if (m < ncols) sts_b32(srec + 96u, __byte_perm(__byte_perm(t[0], t[1], 0x4040), __byte_perm(t[2], t[3], 0x4040), 0x5410));
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          PROF_T(3);
          EPI_BAR();
          PROF_T(4);
          // ---- phase 2: col2im.  Thread = one interior input pixel (r, xo); tap (ky, kx) of phase (ky & 1, kx & 1) reads the
          // partial of pixel (r + offy, xo + offx), offy = (ky + (ky & 1) - 2) / 2
          const int npx = vrows * p.WT;
          const int idx0 = sub * 128 + (warp & 3) * 32 + lane;
          for (int idx = idx0; idx < (DBG(8) ? 0 : npx); idx += 128 * NSUB) {
            const int rr = idx == idx0 ? pre_rr : idx / p.WT, xo = idx == idx0 ? pre_xo : idx - rr * p.WT;
            if (xo >= vcols) continue;
            uint32_t bm[3];
            bm[1] = S_s + (uint32_t)(((rr + 1) * p.P + xo) * DCOL_PIX);
            bm[0] = bm[1] - (uint32_t)(p.P * DCOL_PIX);
            bm[2] = bm[1] + (uint32_t)(p.P * DCOL_PIX);
            uint32_t w[4];
            dcol_gather(bm, w);
#pragma unroll
            for (int ph = 0; ph < 4; ph++) {
              w[ph] = __vadd4(w[ph], dcol_bias);
              w[ph] &= ~prmt_sign_mask(w[ph]);  // ReLU on the wrapped bytes; the pad lane stays 0
            }
            uint8_t* dst = p.out + pm.img_off + ((size_t)(2 * (pm.y0 + rr)) * p.out_x + 2 * (pm.x0 + xo)) * 4;
            *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
            *reinterpret_cast<uint2*>(dst + (size_t)p.out_x * 4) = make_uint2(w[2], w[3]);
          }
          PROF_T(5);
          continue;
        }
        if (DCX) continue;
        if (GEN && DTHIN) {
          // thread = input pixel m of block `half`; its 16 columns are the 2x2 output words it produces (bias + ReLU on the
          // wrapped 8-bit lane, conv_nonsquare_top.cpp:183-194); a warp writes two 256-byte runs of output row 2y and 2y+1
          const int m = half * 128 + q * 32 + lane;
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride + half * 16), v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          const int rr = m / p.P, xo = m - rr * p.P;
          if (xo < vcols && rr < vrows && !DBG(8)) {
            uint32_t w[4];
#pragma unroll
            for (int ph = 0; ph < 4; ph++) {
              uint32_t word = 0;
#pragma unroll
              for (int o = 0; o < 4; o++) {
                if (o < p.OFM) {
                  uint32_t r = (v[ph * 4 + o] + (uint32_t)(int32_t)p.epi.bias[o]) & 0xFFu;
                  r = (r & 0x80u) ? 0u : r;
                  word |= r << (8 * o);
                }
              }
              w[ph] = word;
            }
            uint8_t* dst = p.out + pm.img_off + ((size_t)(2 * (pm.y0 + rr)) * p.out_x + 2 * (pm.x0 + xo)) * 4;
            *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
            *reinterpret_cast<uint2*>(dst + (size_t)p.out_x * 4) = make_uint2(w[2], w[3]);
          }
          continue;
        }
        if (SWPX || (GEN && p.swap)) {
          // Swapped staged bias + ReLU epilogue: thread = pixel (TMEM lane), 32-column loads = 32 channels of that pixel ->
          // 8 packed words -> two 16-byte stores into the SWIZZLE_128B staging row of the pixel; TMA stores un-swizzle.
          const int sb = SWPX ? (half & 1) : alt ? half : (int)ring_idx(acc_it, p.stg_bufs);
          uint8_t* stg = smem + p.stg_off + sb * p.stg_bytes;
          if (lane == 0) {
            if (p.stg_bufs == 2 && !alt) bulk_wait_read<1>();
            else bulk_wait_read<0>();
          }
          if (SWPX || p.wl) __syncwarp();
          else EPI_BAR();
          PROF_T(2);
          // items = (128-pixel block, 32-channel block), block-major; this warp takes every `its`-th item from `it0`
          const int cpb = p.CB * 4, nitems = DBG(8) ? 0 : (p.NPX / 128) * cpb;
          const uint32_t bias_s = smem_u32(smem + p.stage_off);
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
          const uint32_t stg_s = smem_u32(stg);
          const bool fold = p.bias_word >= 0;
          const int mlim = vrows * p.P;
          auto process = [&](int blk, int cbk, const uint32_t (&v)[32]) {
            const int m = blk * 128 + q * 32 + lane;
            if (m >= mlim) return;  // rows of the tile that are never stored
            if (DBG(512)) {  // perf decomposition: TMEM loads only
              if (v[0] == 0x12345678u && v[31] == 0x9abcdef0u) sts_v4(stg_s, v[1], v[2], v[3], v[4]);
              return;
            }
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const uint32_t lo = __byte_perm(v[4 * j], v[4 * j + 1], 0x4040), hi = __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x4040);
              w[j] = __byte_perm(lo, hi, 0x5410);  // low bytes of 4 channels = accumulators mod 256
            }
            if (!fold) {  // a real (warp-uniform) branch: predicated-off byte adds would still cost ~45 issue slots per item
              asm volatile("" ::: "memory");
#pragma unroll
              for (int j = 0; j < 8; j++) w[j] = __vadd4(w[j], lds_b32(bias_s + (uint32_t)(cbk * 32 + 4 * j)));
              asm volatile("" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] &= ~prmt_sign_mask(w[j]);  // ReLU on the wrapped value: bytes with bit 7 set -> 0
            const uint32_t row = stg_s + (uint32_t)((cbk >> 2) * p.NPX * 128 + m * 128);
            const uint32_t c0 = (uint32_t)((cbk & 3) * 2), m7 = (uint32_t)m & 7u;
            sts_v4(row + ((c0 ^ m7) << 4), w[0], w[1], w[2], w[3]);
            sts_v4(row + (((c0 + 1) ^ m7) << 4), w[4], w[5], w[6], w[7]);
          };
          // warp-local mode: after the last channel block of a 128-pixel block this warp's 32 rows are complete
          auto store_block = [&](int blk) {
            fence_proxy_async();
            __syncwarp();
            const int m0 = blk * 128 + q * 32, rr = SWPX ? pre_rr : m0 / p.WT, xo = SWPX ? pre_xo : m0 - rr * p.WT;
            if (lane == 0 && rr < vrows)
              for (int cb = 0; cb < p.CB; cb++) tma_store_4d(&tmO, stg + (cb * p.NPX + m0) * 128, cb * 128, pm.x0 + xo, pm.y0 + rr, img);
          };
          const int it0 = alt ? 0 : half, its = alt ? 1 : 2;
          if (SWPX) {
            // four groups: this group's items are the channel blocks of pixel block half >> 1; one load in flight per warp --
            // 88 registers per thread leave no room for a second buffer, the other three warps of the scheduler cover the latency
            uint32_t va[32];
            const int blk = half >> 1;
#pragma unroll 1
            for (int cbk = 0; cbk < cpb; cbk++) {
              tmem_ld32(tacc + (uint32_t)(blk * p.CB * 128 + cbk * 32), va);
              tmem_ld_wait();
              process(blk, cbk, va);
            }
            store_block(blk);
          } else if (it0 < nitems) {  // software pipeline: the next item's TMEM load is in flight while this one is packed
            uint32_t va[32], vb[32];
            int blk = 0, cbk = it0, left = (nitems - it0 + its - 1) / its;  // items this warp still has to load
            int blk_n = blk, cbk_n = cbk;
            auto advance = [&]() { cbk_n += its; if (cbk_n >= cpb) { cbk_n -= cpb; ++blk_n; } };
            tmem_ld32(tacc + (uint32_t)(blk_n * p.CB * 128 + cbk_n * 32), va);
            --left;
            while (true) {
              tmem_ld_wait();
              PROF_T(3);
              blk = blk_n; cbk = cbk_n;
              if (left > 0) { advance(); tmem_ld32(tacc + (uint32_t)(blk_n * p.CB * 128 + cbk_n * 32), vb); }
              process(blk, cbk, va);
              PROF_T(4);
              if (p.wl && cbk == cpb - 1) { store_block(blk); PROF_T(5); }
              if (left-- <= 0) break;
              tmem_ld_wait();
              PROF_T(3);
              blk = blk_n; cbk = cbk_n;
              if (left > 0) { advance(); tmem_ld32(tacc + (uint32_t)(blk_n * p.CB * 128 + cbk_n * 32), va); }
              process(blk, cbk, vb);
              PROF_T(4);
              if (p.wl && cbk == cpb - 1) { store_block(blk); PROF_T(5); }
              if (left-- <= 0) break;
            }
          }
          PROF_T(7);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          if (SWPX || p.wl) {
            if (lane == 0) bulk_commit();
            PROF_T(6);
            continue;
          }
          fence_proxy_async();
          PROF_T(4);
          EPI_BAR();
          PROF_T(5);
          if (lane == 0) {
            for (int cb = 0; cb < (DBG(8) ? 0 : p.CB); cb++)
              for (int rr = erow0; rr < vrows; rr += erows)
                tma_store_4d(&tmO, stg + (cb * p.NPX + rr * p.P) * 128, cb * 128, pm.x0, pm.y0 + rr, img);
            bulk_commit();
          }
          PROF_T(6);
          continue;
        }
        if (SWPX) continue;
        if (STX || (GEN && p.stg_bufs > 0)) {
          // Staged bias + ReLU epilogue (conv_nonsquare_top.cpp:267-278): thread = channel turns its 32-column TMEM loads into
          // bytes of the tile's [pixel][channel] image in shared memory (a warp's 32 lanes write 32 consecutive bytes: one
          // wavefront, no shuffles, no predicates); one thread then issues a TMA store per tile row.  The register path below
          // (4-byte global stores after a quad transpose) is kept for plans whose planes leave no room for the staging tile.
          const int sb = alt ? (half & 1) : (int)ring_idx(acc_it, p.stg_bufs);
          uint8_t* stg = smem + p.stg_off + sb * p.stg_bytes;
          if (lane == 0) {  // this warp's previous stores from the buffer must have finished reading it
            if (p.stg_bufs == 2 && !alt) bulk_wait_read<1>();
            else bulk_wait_read<0>();
          }
          EPI_BAR();
          PROF_T(2);
          for (int cb = 0; cb < (DBG(8) ? 0 : p.CB); cb++) {
            const int ch = chbase + cb * 128 + q * 32 + lane;
            const uint32_t bias4 = p.bias_word >= 0 ? 0u : ((uint32_t)(int32_t)p.epi.bias[ch] & 0xFFu) * 0x01010101u;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride + cb * p.NPX);
            const uint32_t srow = smem_u32(stg) + (uint32_t)(cb * p.NPX * 128 + q * 32 + lane);
            for (int c0 = col_lo; c0 < col_hi && c0 < vrows * p.P; c0 += col_step) {
              uint32_t v[32];
              tmem_ld32(taddr + (uint32_t)c0, v);
              tmem_ld_wait();
              const uint32_t sa = srow + (uint32_t)c0 * 128u;
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const uint32_t lo = __byte_perm(v[4 * j], v[4 * j + 1], 0x4040), hi = __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x4040);
                uint32_t w = __vadd4(__byte_perm(lo, hi, 0x5410), bias4);  // (acc + bias) mod 256, 4 pixels at once
                w &= ~prmt_sign_mask(w);  // bit 7 set -> 0
                sts_u8(sa + (4 * j + 0) * 128, w);
                sts_u8(sa + (4 * j + 1) * 128, w >> 8);
                sts_u8(sa + (4 * j + 2) * 128, w >> 16);
                sts_u8(sa + (4 * j + 3) * 128, w >> 24);
              }
            }
          }
          PROF_T(3);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);  // one arrival per epilogue warp
          fence_proxy_async();  // generic-proxy writes -> visible to the TMA engine
          PROF_T(4);
          EPI_BAR();
          PROF_T(5);
          if (lane == 0) {  // rows are dealt round-robin to the 8 epilogue warps: issuing a TMA store costs ~300 clocks of one thread
            for (int cb = 0; cb < (DBG(8) ? 0 : p.CB); cb++)
              for (int rr = erow0; rr < vrows; rr += erows) {
                const uint8_t* src = stg + (cb * p.NPX + rr * p.P) * 128;
                if (p.deconv) tma_store_5d(&tmO, src, pm.px * p.OFM + chbase + cb * 128, pm.x0, pm.py, pm.y0 + rr, img);
                else tma_store_4d(&tmO, src, chbase + cb * 128, pm.x0, pm.y0 + rr, img);
              }
            bulk_commit();
          }
          PROF_T(6);
          continue;
        }
        if (STX) continue;
        if (GEN && thin) {
          // Thin output (OFM <= 8, e.g. the 3-channel last layer): only lanes < OFM of the first lane quarter hold data.
          // They turn their 256 columns into bytes in a shared staging row per channel; then all 128 epilogue threads
          // assemble and store whole output words, one pixel per thread (coalesced), instead of 3 lanes doing everything.
          uint8_t* stage = smem + p.stage_off;
          if (q == 0) {  // the whole warp issues the (warp-collective) TMEM loads; lanes >= OFM hold zero rows
            const uint32_t bias4 = (lane < p.OFM ? ((uint32_t)(int32_t)p.epi.bias[lane] & 0xFFu) : 0u) * 0x01010101u;
            const uint32_t taddr = tmem_base + (uint32_t)(acc * p.acc_stride);
            for (int c0 = col_lo; c0 < col_hi && c0 < vrows * p.P; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + (uint32_t)c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const uint32_t lo = __byte_perm(v[4 * j], v[4 * j + 1], 0x4040), hi = __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x4040);
                uint32_t w = __vadd4(__byte_perm(lo, hi, 0x5410), bias4);
                w &= ~prmt_sign_mask(w);
                if (lane < p.OFM) *reinterpret_cast<uint32_t*>(stage + lane * 256 + c0 + 4 * j) = w;
              }
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          for (int m = (warp - 2) * 32 + lane; m < p.NPX; m += 256) {
            const int rr = m / p.P, xo = m - rr * p.P;
            if (xo < vcols && rr < vrows) {
              unsigned long long word = 0;
              for (int c = 0; c < p.OFM; c++) word |= (unsigned long long)stage[c * 256 + m] << (8 * c);
              uint8_t* dst = p.out + pm.word_off(rr, xo, 1);
              if (p.out_word_bytes == 4) *reinterpret_cast<uint32_t*>(dst) = (uint32_t)word;
              else if (p.out_word_bytes == 8) *reinterpret_cast<unsigned long long*>(dst) = word;
              else if (p.out_word_bytes == 2) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)word;
              else *dst = (uint8_t)word;
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");  // staging rows are reused by the next accumulator
        }
        for (int cb = 0; cb < (DBG(8) || thin ? 0 : p.CB); cb++) {
          if (chbase + cb * 128 + q * 32 >= p.OFM) continue;  // this warp's 32 channels do not exist (warp-uniform)
          const int ch = chbase + cb * 128 + q * 32 + lane;
          // threshold tables of this thread's channel: top levels in shared memory + channel-major global row, or all global
          const int chs = chv_of(ch, p.OFM) ? ch : 0;
          const int32_t* tbl = p.epi.thr + chs;
          const int tstride = p.epi.thr_stride;
          const uint32_t top_s = smem_u32(smem + (p.thr_off >= 0 ? p.thr_off : 0)) + 4u * (uint32_t)(cb * 128 + q * 32 + lane);
          const int32_t* row_cm = p.epi.thr_cm + (size_t)chs * (p.epi.thr_n + 1);
          const bool hybrid = p.thr_off >= 0;
          const bool use_lut = p.lut_off >= 0;
          const uint32_t lut_s = smem_u32(smem + (use_lut ? p.lut_off : 0)) + 256u * (uint32_t)(cb * 128 + q * 32 + lane);
          const int32_t lut_lo = cb ? lutlo1 : lutlo0;
          const int lut_sh = cb ? lutsh1 : lutsh0;
          const int row_shift = p.CB == 2 ? 10 : 9;  // log2(CB * 128 channels * 4 bytes)
          const bool chv = ch < p.OFM;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride + cb * p.NPX);
          if (GEN && fast) {
            // bias + ReLU on the wrapped 8-bit value (conv_nonsquare_top.cpp:267-278).  The thread owns one channel and
            // 32 pixels; a 4x4 byte transpose across each lane quad (2 shuffles + 2 PRMT per word) turns that into
            // 4 consecutive channel bytes of one pixel per lane, so a warp store writes 4 pixels x 32 B.
            const uint32_t bias4 = (chv ? ((uint32_t)(int32_t)p.epi.bias[ch] & 0xFFu) : 0u) * 0x01010101u;
            const int li = lane & 3, cgrp = chbase + cb * 128 + q * 32 + (lane & ~3);  // pixel-in-quad, first of this lane's 4 channels
            const uint32_t selA = (lane & 1) ? 0x3715u : 0x6240u, selB = (lane & 2) ? 0x3276u : 0x5410u;
            int rr = (col_lo + li) / p.P, xo = (col_lo + li) - rr * p.P;  // this lane's pixel after the transpose: m = c0 + 4*j + li
            // one running pointer per lane (4 pixels further per word, re-derived at a tile-row wrap) and predicated stores: the
            // epilogue is ~2/3 of this kernel's instructions on MMA-bound layers, and under the power cap instructions are clocks
            const long long xstep = p.deconv ? 2 * p.out_word_bytes : p.out_word_bytes, xstep4 = 4 * xstep;
            const long long wrap_delta = (long long)(p.deconv ? 2 : 1) * p.out_x * p.out_word_bytes - (long long)p.P * xstep;  // next tile row, P pixels back
            uint8_t* dst = p.out + pm.word_off(rr, xo, 1) + cgrp;
            const bool chq = cgrp < p.OFM;
            for (int c0 = col_lo; c0 < col_hi && c0 < vrows * p.P; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + (uint32_t)c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const uint32_t lo = __byte_perm(v[4 * j], v[4 * j + 1], 0x4040), hi = __byte_perm(v[4 * j + 2], v[4 * j + 3], 0x4040);
                uint32_t w = __vadd4(__byte_perm(lo, hi, 0x5410), bias4);      // (acc + bias) mod 256, 4 pixels at once
                w &= ~prmt_sign_mask(w);                                       // bit 7 set -> 0
                const uint32_t a = __byte_perm(w, __shfl_xor_sync(0xffffffffu, w, 1), selA);
                const uint32_t y = __byte_perm(a, __shfl_xor_sync(0xffffffffu, a, 2), selB);
                if (chq && xo < vcols && rr < vrows) *reinterpret_cast<uint32_t*>(dst) = y;
                xo += 4;
                const bool wrapped = xo >= p.P;  // branch-free: a select and two adds
                xo -= wrapped ? p.P : 0;
                rr += wrapped ? 1 : 0;
                dst += xstep4 + (wrapped ? wrap_delta : 0ll);
              }
            }
          } else if (GEN && pk == 1 && p.epi.act_kind == FCB_ACT_THRESHOLDS && p.epi.thr_n == 1 && p.epi.out_bits == 1) {
            // One threshold, 1-bit lanes: the bnn layers on the tensor path (ThresholdsActivation<.., 1, TA, ap_uint<1>>).  The
            // threshold lives in a register, a pixel costs a compare and a ballot, and every lane derives the output word of
            // ITS pixel (column c0 + lane) once per 32-column block instead of once per pixel.
            const int32_t t0 = chv ? __ldg(p.epi.thr + ch) : 0;
            const bool strict = (p.epi.cmp == FCB_CMP_LESS) || (p.epi.cmp == FCB_CMP_GREATER_EQUAL);
            const bool less = (p.epi.cmp == FCB_CMP_LESS) || (p.epi.cmp == FCB_CMP_LESS_EQUAL);
            const int nth = min(p.epi.num_th, 1);  // thr_finish: pos = min(pos, num_th)
            const uint32_t bit0 = (uint32_t)(p.epi.act_val + (less ? 0 : p.epi.num_th)) & 1u;           // no threshold below the accumulator
            const uint32_t bit1 = (uint32_t)(p.epi.act_val + (less ? nth : p.epi.num_th - nth)) & 1u;  // the threshold is below it
            const bool wrap = p.epi.acc_bits < 32;
            const int c0w = chbase + cb * 128 + q * 32;  // first channel of this warp
            int rl = (col_lo + lane) / p.P, xl = (col_lo + lane) - rl * p.P;  // this lane's pixel of the block
#pragma unroll 1
            for (int c0 = col_lo; c0 < col_hi && c0 < vrows * p.P; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + (uint32_t)c0, v);
              tmem_ld_wait();
              uint32_t mine = 0;
#pragma unroll
              for (int j = 0; j < 32; j++) {
                const int32_t a = wrap ? wrap_ta((int32_t)v[j], p.epi.acc_bits, p.epi.acc_signed) : (int32_t)v[j];
                const bool below = strict ? (t0 < a) : (t0 <= a);
                const uint32_t bits = __ballot_sync(0xffffffffu, chv && ((below ? bit1 : bit0) != 0u));
                if (lane == j) mine = bits;
              }
              if (xl < vcols && rl < vrows) {
                uint8_t* dst = p.out + pm.word_off(rl, xl, 1) + (c0w >> 3);
                if (c0w + 32 <= p.OFM && p.out_word_bytes >= 4) *reinterpret_cast<uint32_t*>(dst) = mine;
                else
                  for (int b = 0; b < 4; b++)
                    if (c0w + 8 * b < p.OFM) dst[b] = (uint8_t)(mine >> (8 * b));
              }
              xl += 32;
              while (xl >= p.P) { xl -= p.P; ++rl; }
            }
          } else if (GEN && pk == 1) {
            int rr = col_lo / p.P, xo = col_lo - rr * p.P;
#pragma unroll 1
            for (int c0 = col_lo; c0 < col_hi && rr < vrows; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(taddr + (uint32_t)c0, v);
              tmem_ld_wait();
              uint32_t act[32];
              if (p.epi.act_kind == FCB_ACT_THRESHOLDS) {
#pragma unroll
                for (int b = 0; b < 2; b++) {
                  int32_t a16[16];
                  uint32_t o16[16];
#pragma unroll
                  for (int j = 0; j < 16; j++) a16[j] = (int32_t)v[16 * b + j];
                  if (hybrid) {
#pragma unroll
                    for (int h8 = 0; h8 < 2; h8++) {
                      int32_t a8[8];
                      uint32_t o8[8];
#pragma unroll
                      for (int j = 0; j < 8; j++) a8[j] = a16[8 * h8 + j];
                      if (use_lut) activate_thr_lut<8>(p.epi, top_s, row_shift, lut_s, lut_lo, lut_sh, a8, o8);
                      else activate_thr_hybrid<8>(p.epi, top_s, row_shift, p.thr_top, gshift, row_cm, a8, o8);
#pragma unroll
                      for (int j = 0; j < 8; j++) o16[8 * h8 + j] = o8[j];
                    }
                  } else {
                    activate_thrN<16>(p.epi, tbl, tstride, a16, o16);
                  }
#pragma unroll
                  for (int j = 0; j < 16; j++) act[16 * b + j] = o16[j];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j++) act[j] = chv ? activate(p.epi, ch, (int32_t)v[j]) : 0u;
              }
              if (p.epi.out_bits == 1) {
                // 1-bit lanes: one ballot per pixel gathers the warp's 32 channel bits; lane j keeps pixel j's word so the
                // 32 pixels leave in ONE store instruction (32 single-lane stores into the same few cache lines serialise)
                uint32_t mine = 0;
                size_t mine_off = 0;
                bool mine_ok = false;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                  const uint32_t bits = __ballot_sync(0xffffffffu, chv && (act[j] & 1u));
                  if (lane == j) { mine = bits; mine_ok = xo < vcols && rr < vrows; mine_off = pm.word_off(rr, xo, 1); }
                  if (++xo == p.P) { xo = 0; ++rr; }
                }
                const int c0 = chbase + cb * 128 + q * 32;  // first channel of this warp
                if (mine_ok) {
                  uint8_t* dst = p.out + mine_off + (c0 >> 3);
                  if (c0 + 32 <= p.OFM && p.out_word_bytes >= 4) *reinterpret_cast<uint32_t*>(dst) = mine;
                  else
                    for (int b = 0; b < 4; b++)
                      if (c0 + 8 * b < p.OFM) dst[b] = (uint8_t)(mine >> (8 * b));
                }
              } else {
                if (p.epi.out_bits == 8) {
                  // byte lanes: one pointer per thread, advanced pixel by pixel (a warp store = 32 consecutive bytes of one word)
                  uint8_t* dst = p.out + pm.word_off(rr, xo, 1) + ch;
                  const int xstep = p.deconv ? 2 * p.out_word_bytes : p.out_word_bytes;  // a phase's pixels are every other output word
#pragma unroll
                  for (int j = 0; j < 32; j++) {
                    if (xo < vcols && rr < vrows && chv) *dst = (uint8_t)act[j];
                    dst += xstep;
                    if (++xo == p.P) { xo = 0; ++rr; dst = p.out + pm.word_off(rr, 0, 1) + ch; }
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; j++) {
                    if (xo < vcols && rr < vrows) store_lane(p.out + pm.word_off(rr, xo, 1), ch, chv, act[j], p.epi.out_bits);  // warp-uniform
                    if (++xo == p.P) { xo = 0; ++rr; }
                  }
                }
              }
            }
          } else {
            // 2x2 max pool: rows rr, rr+1 of the tile are columns m and m + P of the same thread
            // work units = (row pair, 16-column block) = 8 pooled outputs; the two warps of a lane quarter take alternate units
            const int xblocks = (vcols + 15) >> 4, nunits = (vrows >> 1) * xblocks;
            int urow = 0, ublk = half;  // unit -> (row pair, x block), advanced without divisions (half < NHALF <= 4)
            while (ublk >= xblocks) { ublk -= xblocks; urow += 2; }
#pragma unroll 1
            for (int u = half; u < nunits; u += NHALF) {
              {
                const int rr = urow, xb = 16 * ublk;
                ublk += NHALF;
                while (ublk >= xblocks) { ublk -= xblocks; urow += 2; }
                if (mono) {
                  // count-of-thresholds-below is non-decreasing in the TA-wrapped accumulator: pool first, then ONE search
                  int32_t m8[8];
                  uint32_t pooled[8];
                  uint32_t va[2][8] = {}, vb[2][8] = {};
                  const bool second = xb + 8 < vcols;  // warp-uniform; 8-column groups keep reads inside the accumulator stage
                  tmem_ld8(taddr + (uint32_t)(rr * p.P + xb), va[0]);
                  tmem_ld8(taddr + (uint32_t)((rr + 1) * p.P + xb), vb[0]);
                  if (second) {
                    tmem_ld8(taddr + (uint32_t)(rr * p.P + xb + 8), va[1]);
                    tmem_ld8(taddr + (uint32_t)((rr + 1) * p.P + xb + 8), vb[1]);
                  }
                  tmem_ld_wait();
                  if (p.epi.acc_bits >= 32) {
                    // the create-time range analysis found the sums inside TA (fcb_api.cu): no wrap, three max instructions per output
                    // (one uniform branch here instead of predicated shifts on every accumulator)
#pragma unroll
                    for (int g2 = 0; g2 < 2; g2++) {
#pragma unroll
                      for (int w = 0; w < 4; w++) {
                        const int32_t m0 = max(max((int32_t)va[g2][2 * w], (int32_t)va[g2][2 * w + 1]), max((int32_t)vb[g2][2 * w], (int32_t)vb[g2][2 * w + 1]));
                        m8[4 * g2 + w] = (g2 == 0 || second) ? m0 : 0;
                      }
                    }
                  } else {
#pragma unroll
                    for (int g2 = 0; g2 < 2; g2++) {
#pragma unroll
                      for (int w = 0; w < 4; w++) {
                        int32_t m0 = wrap_ta((int32_t)va[g2][2 * w], p.epi.acc_bits, p.epi.acc_signed);
                        m0 = max(m0, wrap_ta((int32_t)va[g2][2 * w + 1], p.epi.acc_bits, p.epi.acc_signed));
                        m0 = max(m0, wrap_ta((int32_t)vb[g2][2 * w], p.epi.acc_bits, p.epi.acc_signed));
                        m0 = max(m0, wrap_ta((int32_t)vb[g2][2 * w + 1], p.epi.acc_bits, p.epi.acc_signed));
                        m8[4 * g2 + w] = (g2 == 0 || second) ? m0 : 0;
                      }
                    }
                  }
                  if (DBG(16)) {
#pragma unroll
                    for (int w = 0; w < 8; w++) pooled[w] = (uint32_t)m8[w] & 0xFFu;
                  } else if (THRP && use_lut) {
                    // (m8 is already wrapped to TA; tables in shared memory carry the less_equal adjustment)
                    const uint32_t rowb = p.CB == 2 ? 1024u : 512u, tbase = top_s - rowb, fbase = tbase - (uint32_t)p.epi.act_val * rowb;
                    if (p.CB == 2) {
                      if (p.epi.thr_lut_levels == 3) thr_lut_fast<8, 3, 1024>(tbase, fbase, lut_s, lut_lo, lut_sh, m8, pooled);
                      else thr_lut_fast<8, 4, 1024>(tbase, fbase, lut_s, lut_lo, lut_sh, m8, pooled);
                    } else {
                      if (p.epi.thr_lut_levels == 3) thr_lut_fast<8, 3, 512>(tbase, fbase, lut_s, lut_lo, lut_sh, m8, pooled);
                      else thr_lut_fast<8, 4, 512>(tbase, fbase, lut_s, lut_lo, lut_sh, m8, pooled);
                    }
                  } else if (use_lut) {
                    activate_thr_lut<8>(p.epi, top_s, row_shift, lut_s, lut_lo, lut_sh, m8, pooled);
                  } else if (THRP || hybrid) {
                    activate_thr_hybrid<8, false>(p.epi, top_s, row_shift, p.thr_top, gshift, row_cm, m8, pooled);  // (m8 is TA-wrapped above)
                  } else {
                    activate_thrN<8>(p.epi, tbl, tstride, m8, pooled);
                  }
                  if (THRP || p.epi.out_bits == 8) {
                    // x0 and xb are even: pooled pixel w sits w words further.  A warp store = 32 consecutive bytes of one word.
                    uint8_t* dst = p.out + (THRP ? pm.pooled2_off(rr, xb) : pm.word_off(rr, xb, 2)) + ch;
                    const int owb = p.out_word_bytes;
                    if (chv && xb + 16 <= vcols) {  // whole unit (warp-uniform except the channel tail): no per-store predicates,
                      if (owb == 128) {              // immediate offsets for the common word sizes
#pragma unroll
                        for (int w = 0; w < 8; w++) dst[w * 128] = (uint8_t)pooled[w];
                      } else if (owb == 256) {
#pragma unroll
                        for (int w = 0; w < 8; w++) dst[w * 256] = (uint8_t)pooled[w];
                      } else {
#pragma unroll
                        for (int w = 0; w < 8; w++) dst[(unsigned)(w * owb)] = (uint8_t)pooled[w];
                      }
                    } else if (chv) {
#pragma unroll
                      for (int w = 0; w < 8; w++)
                        if (xb + 2 * w < vcols) dst[(unsigned)(w * owb)] = (uint8_t)pooled[w];
                    }
                  } else {
#pragma unroll
                    for (int w = 0; w < 8; w++) {
                      const int xo = xb + 2 * w;
                      if (xo < vcols) store_lane(p.out + pm.word_off(rr, xo, 2), ch, chv, pooled[w], p.epi.out_bits);  // whole windows
                    }
                  }
                } else {
#pragma unroll 1
                  for (int g4 = 0; g4 < 2; g4++) {
                    uint32_t va[8], vb[8];
                    if (xb + 8 * g4 >= vcols) break;
                    tmem_ld8(taddr + (uint32_t)(rr * p.P + xb + 8 * g4), va);
                    tmem_ld8(taddr + (uint32_t)((rr + 1) * p.P + xb + 8 * g4), vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                      uint32_t a = 0;
                      if (chv) {
                        a = activate(p.epi, ch, (int32_t)va[2 * w]);
                        a = max(a, activate(p.epi, ch, (int32_t)va[2 * w + 1]));
                        a = max(a, activate(p.epi, ch, (int32_t)vb[2 * w]));
                        a = max(a, activate(p.epi, ch, (int32_t)vb[2 * w + 1]));
                      }
                      const int xo = xb + 8 * g4 + 2 * w;
                      if (xo < vcols) store_lane(p.out + pm.word_off(rr, xo, 2), ch, chv, a, p.epi.out_bits);
                    }
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
    }
  }
  if (warp == 2 && lane == 0) PROF_FLUSH(2);
  if (warp >= 2 && (warp < 10 || warp >= 10 + NB) && lane == 0 && p.stg_bufs > 0) bulk_wait<0>();  // staged stores still read shared memory
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
static int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct TapH { int offx, offy, parx, pary, wtap, phase; };

int umma2_plan_create(const Geom& g, const int8_t* d_w, const EpiParams& epi, int num_sms, Umma2Plan** out) {
  *out = nullptr;
  if (g.pool > 2) return FCB_ERR_UNSUPPORTED;
  const int deconv = g.kind == FCB_KIND_DECONV522;
  if (deconv && g.pool != 1) return FCB_ERR_UNSUPPORTED;
  const int s = deconv ? 1 : g.SX;
  const int cch = (g.C + 127) / 128;  // a partial last chunk is zero-filled by TMA (stride-1 views only)
  const int CB = (g.OFM + 127) / 128;
  if (CB > 2) return FCB_ERR_UNSUPPORTED;
  // ---- taps
  std::vector<TapH> taps;
  const int nph = deconv ? 4 : 1;
  if (!deconv) {
    for (int ky = 0; ky < g.KY; ky++)
      for (int kx = 0; kx < g.KX; kx++) {
        // FMPadding_nonsquare's left / up share (streamtools.h:374-379): the right / down zeros are TMA out-of-bounds fill like the
        // rest, so an asymmetric split only moves the tap offsets (a dilated tap is just another plane offset as well)
        const int dx = kx * g.DX - g.pad_l, dy = ky * g.DY - g.pad_u;
        taps.push_back({fdiv(dx, s), fdiv(dy, s), dx - fdiv(dx, s) * s, dy - fdiv(dy, s) * s, ky * g.KX + kx, 0});
      }
  } else {
    for (int py = 0; py < 2; py++)
      for (int px = 0; px < 2; px++)
        for (int ky = 0; ky < 5; ky++)
          for (int kx = 0; kx < 5; kx++) {
            if (((px + kx) & 1) || ((py + ky) & 1)) continue;  // structural zero
            taps.push_back({(px + kx - 2) / 2, (py + ky - 2) / 2, 0, 0, ky * 5 + kx, py * 2 + px});
          }
  }
  // ---- planes: one per (pary, parx, channel chunk) that has taps
  struct PlaneH { int pary, parx, minx, maxx, miny, maxy, used; };
  PlaneH ph4[4];
  for (int i = 0; i < 4; i++) ph4[i] = {i / 2, i % 2, 1 << 20, -(1 << 20), 1 << 20, -(1 << 20), 0};
  for (auto& t : taps) {
    PlaneH& h = ph4[t.pary * 2 + t.parx];
    h.used = 1;
    h.minx = std::min(h.minx, t.offx); h.maxx = std::max(h.maxx, t.offx);
    h.miny = std::min(h.miny, t.offy); h.maxy = std::max(h.maxy, t.offy);
  }
  int nparity = 0, halo_x = 0;
  for (auto& h : ph4) if (h.used) { nparity++; halo_x = std::max(halo_x, h.maxx - h.minx); }
  if (nparity * cch > U2_MAX_PLANES) return FCB_ERR_UNSUPPORTED;
  int max_kb = 0;
  for (int p = 0; p < nph; p++) {
    int n = 0;
    for (auto& t : taps) n += (t.phase == p);
    max_kb = std::max(max_kb, n * cch);
  }
  if (max_kb > U2_MAX_KB) return FCB_ERR_UNSUPPORTED;
  const int PX = deconv ? g.IX : g.OX, PY = deconv ? g.IY : g.OY;
  const int xdim = s == 2 ? g.IX / 2 : g.IX, ydim = s == 2 ? g.IY / 2 : g.IY;  // tensor extents the TMA box must fit

  // ---- choose (WT, R, NPX, wstages): minimise estimated clocks per useful output pixel.
  // Threshold layers keep the top levels of the per-channel binary search in shared memory (activate_thr_hybrid):
  // all D levels if they fit ~66 KB, else D-2 (the rest is one 16-byte global load per output), else none.
  const int step = g.pool == 2 ? 2 : 1;
  const bool thr = g.act_kind == FCB_ACT_THRESHOLDS;
  // Candidates for where the per-channel threshold tables live (activate_thr_hybrid): the whole search in shared memory
  // (D levels), the whole search with the channel blocks split over CTA groups (chb = CB: each CTA keeps 128 channels'
  // tables and visits every tile), the top D-2 levels (+ one 16-byte global load per output), or global memory only.
  struct Cand { int thr_top, chb; double epi_factor; };
  std::vector<Cand> cands;
  if (thr && !exp_env("FCB_U2_NO_SMEM_THR")) {
    int D = 0;
    while ((1 << D) < epi.thr_n + 1) D++;
    cands.push_back({D, 1, 40.0});
    if (CB == 2 && !exp_env("FCB_U2_NO_CHB")) cands.push_back({D, 2, 40.0});
    if (D - 2 >= 1) cands.push_back({D - 2, 1, 60.0});
  }
  cands.push_back({0, 1, thr ? 120.0 : 0.0});
  // two channel blocks without shared-memory tables: one CTA group per block keeps 256 accumulator columns free, so the epilogue of
  // a tile overlaps the MMAs of the next (CB = 2 in one CTA fills all 512 columns: acc_stages = 1, epilogue serial with the MMAs)
  if (!thr && CB == 2 && !exp_env("FCB_U2_NO_CHB")) cands.push_back({0, 2, 0.0});
  // two channel blocks, hybrid tables per CTA group (top D-2 levels of 128 channels: 32 KB, room for two plane sets and N = 256).
  // The factor is calibrated, not derived: config 5b stage 4 473 k -> 530 k img/s against the whole-table chb = 2 form, config 4
  // 430 k -> 504 k against the D-2 hybrid on N = 128 tiles (profiles/r02_thr_candidates.log), so it must win against both (2 x 36 < 2 x 40).
  if (thr && CB == 2 && !exp_env("FCB_U2_NO_SMEM_THR") && !exp_env("FCB_U2_NO_CHB")) {
    int D = 0;
    while ((1 << D) < epi.thr_n + 1) D++;
    if (D - 2 >= 1 && D >= 7 && exp_int("FCB_U2_CHB_HYB", 1)) cands.push_back({D - 2, 2, 36.0});
  }
  // Big tables (>= 127 thresholds) on one channel block: the whole table in shared memory is 8 plain levels, leaves room for one plane
  // set only, and loses to the D-2 hybrid with double-buffered planes (config 5b stage 2 / 3: 65.9 k / 254 k vs 71.4 k / 279 k img/s,
  // profiles/r02_thr_candidates.log).  Making room for the bucket LUT beside the table (N = 128 tiles, two weight stages) was measured
  // too: 52 k / 187 k img/s -- the short tiles cost more than the cheaper search returns.
  int Dfull = 0;
  if (thr) while ((1 << Dfull) < epi.thr_n + 1) Dfull++;
  bool best_lut = false;
  double best = 1e30;
  int bWT = 0, bR = 0, bNPX = 0, bWS = 0, thr_top = 0, thr_bytes = 0, chb = 1, CBe = CB;
  const int only_cand = exp_int("FCB_U2_CAND", -1);  // experiments: evaluate one candidate only
  for (size_t ci = 0; ci < cands.size(); ci++) {
    const Cand& cd = cands[ci];
    if (only_cand >= 0 && (int)ci != only_cand) continue;
    const int cbe = CB / cd.chb;
    const int tb = cd.thr_top ? ((1 << cd.thr_top) - 1) * cbe * 128 * 4 : 0;
    if (tb > 136 * 1024) continue;
    const int smem_limit = 227 * 1024 - 2048 - 2048 - (tb ? tb + 128 : 0);
    const int w_bytes_c = cbe * 128 * 128;
    for (int NPX = 256; NPX >= 64; NPX /= 2) {
      if (cbe * NPX > 512) continue;
      const int acc_st = (2 * cbe * NPX <= 512) ? 2 : 1;
      for (int WS = 4; WS >= 2; WS--)
        for (int WT = step; WT <= std::min(PX + step - 1, 254 - halo_x); WT += step) {
          const int P = WT + halo_x;
          if (P > xdim) continue;
          const int budget = (g.pool == 2 && (P % 8)) ? NPX - 8 : NPX;  // pooled epilogue reads 8-column groups from row starts
          int R = std::min(budget / P, PY);
          if (g.pool == 2) R &= ~1;
          if (R < 1) continue;
          size_t plane_bytes = 0;
          bool ok = true;
          for (auto& h : ph4) {
            if (!h.used) continue;
            const int rows = R + (h.maxy - h.miny);
            if (rows > 256 || rows > ydim) { ok = false; break; }
            const int a_off_max = (h.maxy - h.miny) * P + (h.maxx - h.minx);
            size_t b = std::max<size_t>((size_t)rows * P * 128, (size_t)(a_off_max + NPX) * 128);
            plane_bytes += (b + 1023) / 1024 * 1024 * cch;
          }
          if (!ok || (long long)plane_bytes + (long long)WS * w_bytes_c > (long long)smem_limit) continue;
          const double tiles = (double)((PX + WT - 1) / WT) * ((PY + R - 1) / R) * cd.chb;  // tile visits per CTA group
          const double nkb = (double)taps.size() * cch;  // all phases
          // clocks per instruction: measured 146 at N = 256 and 102 at N = 128 = shared-memory operand reads at ~84 B/clk
          // (4 KB of weights + 32 B per column); N is trimmed to the tile's last useful column below
          const int n_eff = std::min(NPX, ((R - 1) * P + WT + 15) / 16 * 16);
          const double mma_clk = nkb * cbe * 4 * (58.0 + 0.344 * n_eff);
          const double fill_clk = ((double)plane_bytes + nkb * w_bytes_c) / 38.0;  // measured L2->SM fill, B/clk/SM
          // epilogue: bias/ReLU ~7 clk per pixel and channel block.  Threshold search: instruction bound, epi_factor issue clocks
          // per warp-level output (32 channels), counted with the padding of the 16-wide search batches
          // (profiles/r01_cfg4_epilogue_profile.txt)
          const double groups = g.pool == 2 ? (double)(R / 2) * ((WT + 15) / 16) * 8 : (double)((R * P + 31) / 32) * 32;
          double ef = cd.epi_factor;
          const bool with_lut = false;
          if (thr && cd.thr_top == Dfull && Dfull >= 7 && cd.chb == 1) ef = 66.0;
          const double epi_clk = thr ? groups * nph * cbe * 4 * ef : (double)WT * R * nph * cbe * 7.0;
          const double tile_clk = acc_st == 2 ? std::max(std::max(mma_clk, fill_clk), epi_clk) : std::max(mma_clk + epi_clk, fill_clk);
          const double ws_pen = WS >= 3 ? 1.0 : 1.05;
          // ties (e.g. 1x1 layers, where every WT is equally efficient) go to wide boxes: long contiguous TMA rows
          const double cost = tiles * tile_clk * ws_pen * (1.0 + 0.0005 * R) / ((double)PX * PY);
          if (cost < best * 0.999) {
            best = cost; bWT = WT; bR = R; bNPX = NPX; bWS = WS;
            thr_top = cd.thr_top; thr_bytes = tb; chb = cd.chb; CBe = cbe; best_lut = with_lut;
          }
        }
    }
  }
  const int w_bytes = CBe * 128 * 128;
  if (const char* force = exp_env("FCB_U2_FORCE")) {  // "WT,R,NPX,WS" -- experiments only; the caller is responsible for it fitting
    int a, b, c, d;
    if (sscanf(force, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) { bWT = a; bR = b; bNPX = c; bWS = d; }
  }
  if (!bWT) return FCB_ERR_UNSUPPORTED;

  std::unique_ptr<Umma2Plan> U_owner(new Umma2Plan());  // released to the caller on success; every early return frees it
  Umma2Plan* U = U_owner.get();
  U->g = g; U->num_sms = num_sms;
  Params2& p = U->p;
  memset(&p, 0, sizeof(p));
  p.epi = epi;
  p.OFM = g.OFM; p.CB = CBe; p.chb = chb; p.NPX = bNPX; p.stride2 = (s == 2); p.deconv = deconv; p.nphases = nph;
  p.WT = bWT; p.R = bR; p.P = bWT + halo_x; p.PX = PX; p.PY = PY;
  p.tiles_x = (PX + bWT - 1) / bWT; p.tiles_y = (PY + bR - 1) / bR;
  p.out_x = g.out_x; p.out_y = g.out_y; p.out_word_bytes = (int)g.out_word_bytes; p.out_img_bytes = g.out_img_bytes;
  p.wstages = bWS; p.w_bytes = w_bytes; p.ksteps = 4; p.bias_word = -1;
  p.acc_stride = CBe * bNPX;
  p.acc_stages = (2 * p.acc_stride <= 512) ? 2 : 1;
  int tc = 32;
  while (tc < p.acc_stages * p.acc_stride) tc *= 2;
  p.tmem_cols = tc;
  // MMA width: the last accumulator column any output of the tile maps to is (R-1)*P + WT - 1; columns past it (the tail of the 256)
  // are never read back as results, so N stops at the next multiple of 16 (an MMA costs ~(4 KB of A + 32 B x N of B) / 84 B/clk)
  int n_mma = bNPX;
  if (exp_int("FCB_U2_NTRIM", 1)) n_mma = std::min(bNPX, ((bR - 1) * (bWT + halo_x) + bWT + 15) / 16 * 16);
  p.idesc = make_idesc_i8(128, n_mma, /*A = weights*/ 1, /*B = activations*/ g.in_signed);
  p.n_mma = n_mma;
  p.debug = exp_int("FCB_U2_DEBUG", 0);
  // planes
  int plane_of[4][4];
  int off = 0, np = 0, nmaps = 0;
  int map_rows[2] = {0, 0};
  for (int i = 0; i < 4; i++) {
    PlaneH& h = ph4[i];
    if (!h.used) continue;
    const int rows = bR + (h.maxy - h.miny);
    int map = -1;
    for (int m = 0; m < nmaps; m++) if (map_rows[m] == rows) map = m;
    if (map < 0) {
      if (nmaps == 2) return FCB_ERR_UNSUPPORTED;
      map = nmaps; map_rows[nmaps++] = rows;
    }
    const int a_off_max = (h.maxy - h.miny) * p.P + (h.maxx - h.minx);
    size_t b = std::max<size_t>((size_t)rows * p.P * 128, (size_t)(a_off_max + bNPX) * 128);
    b = (b + 1023) / 1024 * 1024;
    for (int cc = 0; cc < cch; cc++) {
      Plane2& pl = p.planes[np];
      pl.smem_off = off; pl.bytes = rows * p.P * 128;
      pl.c0 = (p.stride2 ? h.parx * g.C : 0) + cc * 128;
      pl.dx = h.minx; pl.dy = h.miny; pl.par = h.pary; pl.map = map;
      plane_of[i][cc] = np++;
      off += (int)b;
    }
  }
  p.nplanes = np;
  p.set_bytes = off;
  const size_t lut_reserve = best_lut ? (size_t)CBe * 128 * 256 + 128 : 0;  // (the LUT the cost model counted on comes before a second plane set)
  p.nsets = ((size_t)2 * off + (size_t)bWS * w_bytes + 4096 + (thr_bytes ? thr_bytes + 128 : 0) + lut_reserve <= (size_t)227 * 1024 - 1024) ? 2 : 1;
  p.nsets = std::max(1, std::min(p.nsets, exp_int("FCB_U2_NSETS", p.nsets)));
  off *= p.nsets;
  U->box_rows[0] = map_rows[0]; U->box_rows[1] = nmaps > 1 ? map_rows[1] : map_rows[0];
  p.w_off = off;
  off += bWS * w_bytes;
  p.bar_off = off;
  off += (2 * bWS + 2 * np * p.nsets + 4 + 1 + 2 * (int)U2_NPB + 1) * 8;
  off = (off + 15) & ~15;
  p.stage_off = off;
  off += 8 * 256;
  p.thr_off = -1;
  p.thr_top = thr_top;
  if (thr_bytes) {
    off = (off + 127) & ~127;
    p.thr_off = off;
    off += thr_bytes;
  }
  p.lut_off = -1;
  {
    int D = 0;
    while ((1 << D) < epi.thr_n + 1) D++;
    const int lut_bytes = CBe * 128 * 256;
    if (thr_bytes && thr_top == D && epi.thr_lut && !exp_env("FCB_U2_NO_LUT") && (size_t)off + lut_bytes + 2048 <= (size_t)227 * 1024) {
      off = (off + 127) & ~127;
      p.lut_off = off;
      off += lut_bytes;
    }
  }
  // staged epilogue (bias + ReLU on whole 128-channel blocks): two buffers if they fit, else one, else the register path
  p.stg_off = 0; p.stg_bufs = 0; p.stg_bytes = CBe * bNPX * 128;
  const bool fast_epi = epi.act_kind == FCB_ACT_BIAS_RELU && epi.out_bits == 8 && epi.acc_bits == 8 && g.pool <= 1 && g.OFM % 128 == 0 &&
                        g.out_word_bytes == (size_t)g.OFM;
  if (fast_epi && !exp_env("FCB_U2_NO_STAGE")) {
    off = (off + 127) & ~127;
    for (int nb = 2; nb >= 1; nb--)
      if ((size_t)off + (size_t)nb * p.stg_bytes + 1024 <= (size_t)227 * 1024) { p.stg_bufs = nb; break; }
    p.stg_off = off;
    off += p.stg_bufs * p.stg_bytes;
  }
  p.epi_alt = (p.stg_bufs == 2 && p.acc_stages == 2 && !exp_env("FCB_U2_NO_ALT")) ? 1 : 0;
  {
    // register bias/ReLU epilogue, ~30 clocks per pixel with 4 warps (measured on the deconv layers before they were staged):
    // if that hides under the tile's MMA time with margin, run it on 4 warps
    const bool reg_fast = epi.act_kind == FCB_ACT_BIAS_RELU && epi.out_bits == 8 && epi.acc_bits == 8 && g.pool <= 1 && g.OFM % 32 == 0 &&
                          p.P >= 4 && p.stg_bufs == 0 && g.OFM > 8;
    const double mma_tile = (double)taps.size() * cch * CBe * 4 * (bNPX == 256 ? 146.0 : 102.0) / nph;
    const double epi_tile = (double)bWT * bR * CBe * 30.0;
    // measured on CONV_1: 147.3 k img/s with 4 warps vs 150.5 k with 8 -- issue-slot competition is not what holds the tensor
    // pipe at ~85 %, so this stays an experiment switch (FCB_U2_EPI4=1)
    (void)mma_tile; (void)epi_tile;
    p.epi4 = (reg_fast && p.acc_stages == 2 && exp_int("FCB_U2_EPI4", 0)) ? 1 : 0;
  }
  U->smem = (size_t)off + 1024;
  if (U->smem > 227 * 1024) return FCB_ERR_UNSUPPORTED;
  // K-block lists: plane-major within each phase so planes are released progressively
  std::vector<int> last_use(np, -1), first_use(np, -1);
  int ord = 0;
  for (int phs = 0; phs < nph; phs++) {
    Phase2& P = p.phases[phs];
    P.nkb = 0; P.px = phs % 2; P.py = phs / 2;
    for (int i = 0; i < 4; i++) {
      if (!ph4[i].used) continue;
      for (int cc = 0; cc < cch; cc++)
        for (auto& t : taps) {
          if (t.phase != phs || t.pary * 2 + t.parx != i) continue;
          KB2& kb = P.kb[P.nkb++];
          kb.plane = (uint16_t)plane_of[i][cc];
          kb.flags = 0;
          kb.a_off = (t.offy - ph4[i].miny) * p.P + (t.offx - ph4[i].minx);
          kb.w_k = t.wtap * g.C + cc * 128;
          kb.d_off = (uint32_t)(p.planes[kb.plane].smem_off + kb.a_off * 128) >> 4;
          if (first_use[kb.plane] < 0) first_use[kb.plane] = ord;
          last_use[kb.plane] = ord;
          ord++;
        }
    }
  }
  ord = 0;
  for (int phs = 0; phs < nph; phs++)
    for (int i = 0; i < p.phases[phs].nkb; i++, ord++) {
      KB2& kb = p.phases[phs].kb[i];
      if (first_use[kb.plane] == ord) kb.flags |= KB_WAIT;
      if (last_use[kb.plane] == ord) kb.flags |= KB_FREE;
    }
  // weights as the A operand: box = 128 B of K x CB*128 channel rows; rows >= OFM are zero-filled by TMA
  {
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)(CB * 128)};  // rows >= OFM are zeros in memory (fcb_umma.cu)
    const uint64_t strides[1] = {(uint64_t)g.K};
    const uint32_t box[2] = {128, (uint32_t)(CBe * 128)};
    int rc = umma_encode_map(&U->tmW, const_cast<int8_t*>(d_w), 2, dims, strides, box);
    if (rc) return rc;
    // 2-CTA clusters share the weight stream: each CTA fetches half of the rows of a K-block (a multiple of 8 rows, so the slice
    // starts on a 1024-byte swizzle period) and multicasts it.  Not with chb > 1: neighbouring CTAs then serve different channel blocks.
    // MEASURED (profiles/r02_cluster_multicast_ab.log, sustained, power-capped): the halved L2 stream buys ~1.5 % of SM clock, the
    // lock-step of the two CTAs on one ring costs ~3 % of clocks per image (CONV_1 119.5 k vs 121.1 k img/s, L6 113.6 k vs 115.6 k):
    // a net loss, so the clustered instantiations are compiled into experiment builds only (FCB_U2_CLUSTER=1).
    U->cluster_ok = chb == 1 && exp_int("FCB_U2_CLUSTER", 0) != 0;
    if (U->cluster_ok) {
      const uint32_t box2[2] = {128, (uint32_t)(CBe * 64)};
      rc = umma_encode_map(&U->tmW_slice, const_cast<int8_t*>(d_w), 2, dims, strides, box2);
      if (rc) return rc;
    }
  }
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#ifdef FCB_EXPERIMENT
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 0, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#endif
  *out = U_owner.release();
  return FCB_OK;
}

// launch of a resident-planes instantiation: as 2-CTA clusters sharing the weight stream when the plan allows it
template <int EPI>
static cudaError_t launch_resident(const Umma2Plan* U, int grid, int threads, cudaStream_t st, const CUtensorMap& a0, const CUtensorMap& a1,
                                   const CUtensorMap& o, const Params2& p) {
#ifdef FCB_EXPERIMENT
  if (U->cluster_ok && grid >= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(grid & ~1), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = U->smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, umma2_conv_kernel<1, 0, EPI, 2>, a0, a1, U->tmW_slice, o, p);
  }
#endif
  umma2_conv_kernel<1, 0, EPI, 1><<<grid, threads, U->smem, st>>>(a0, a1, U->tmW, o, p);
  return cudaGetLastError();
}

// Thin-input plan: d_w is [CB*128][128] s8 with k = (ky*KX + kx)*4 + lane (zero beyond the window / lanes >= C).
int umma2_plan_create_thin(const Geom& g, const int8_t* d_w, const EpiParams& epi, int num_sms, int bias_word, Umma2Plan** out) {
  *out = nullptr;
  const int s = g.SX, nw = g.KX * g.KY;
  if (g.kind != FCB_KIND_CONV || g.in_word_bytes != 4 || g.in_bits != 8 || nw > 32 || g.SX != g.SY || s < 1 || s > 2 || g.pool > 2 || g.DX != 1 || g.DY != 1 ||
      (g.IX % 4) || g.OFM > 256)
    return FCB_ERR_UNSUPPORTED;
  const int CB = (g.OFM + 127) / 128;
  const int PX = g.OX, PY = g.OY;
  const int step = g.pool == 2 ? 2 : 1;
  const bool thr = g.act_kind == FCB_ACT_THRESHOLDS;
  int thr_top = 0, thr_bytes = 0;
  if (thr && !exp_env("FCB_U2_NO_SMEM_THR")) {
    int D = 0;
    while ((1 << D) < epi.thr_n + 1) D++;
    // the whole search in shared memory when it fits beside the small operands of this mode, else all but the last two levels
    const int cand[2] = {D, D - 2};
    for (int c : cand) {
      if (c < 1) continue;
      const int bytes = ((1 << c) - 1) * CB * 128 * 4;
      if (bytes <= 132 * 1024) { thr_top = c; thr_bytes = bytes; break; }
    }
  }
  const int w_bytes = CB * 128 * 128;
  const int ksteps = ((bias_word >= 0 ? nw + 1 : nw) * 4 + 31) / 32;
  int bWT = 0, bR = 0, bNPX = 0;
  double best = 1e30;
  const bool fast_epi0 = epi.act_kind == FCB_ACT_BIAS_RELU && epi.out_bits == 8 && epi.acc_bits == 8 && g.pool <= 1 && g.OFM % 128 == 0 &&
                         g.out_word_bytes == (size_t)g.OFM && !exp_env("FCB_U2_NO_STAGE");
  const bool want_swap = fast_epi0 && PX >= 8 && !exp_env("FCB_U2_NO_SWAP");
  for (int NPX = 256; NPX >= (want_swap ? 128 : 64); NPX /= 2) {
    if (CB * NPX > 512) continue;
    for (int WT = step; WT <= std::min(PX + step - 1, 256); WT += step) {
      // swizzled staging rows: every tile row starts on a 1024-byte boundary; rows of whole 32-pixel segments allow warp-local stores
      if (want_swap && (WT % (PX >= 32 ? 32 : 8))) continue;
      const int BWp = (s * (WT - 1) + g.KX + 3 + 3) / 4 * 4;  // + up to 3 pixels of alignment slack on the left
      if (BWp > 256) continue;
      const int budget = (g.pool == 2 && (WT % 8)) ? NPX - 8 : NPX;  // pooled epilogue reads 8-column groups from row starts
      int R = std::min(budget / WT, PY);
      if (g.pool == 2) R &= ~1;
      if (R < 1) continue;
      const int BHp = s * (R - 1) + g.KY;
      if (BHp > 256) continue;
      const size_t need = (size_t)2 * NPX * 128 + w_bytes + U2_NPB * (((size_t)BWp * BHp * 4 + 127) / 128 * 128) + 8192 + (thr_bytes ? thr_bytes + 128 : 0);
      if (need > (size_t)227 * 1024) continue;
      const double tiles = (double)((PX + WT - 1) / WT) * ((PY + R - 1) / R);
      // per tile: fixed cost ~ NPX columns of MMA/epilogue work + overheads; prefer full tiles and wide rows
      const double cost = tiles * (NPX + 24.0) * (1.0 + 0.002 * R) / ((double)PX * PY);
      if (cost < best * 0.999) { best = cost; bWT = WT; bR = R; bNPX = NPX; }
    }
  }
  if (!bWT) return FCB_ERR_UNSUPPORTED;
  std::unique_ptr<Umma2Plan> U_owner(new Umma2Plan());  // released to the caller on success; every early return frees it
  Umma2Plan* U = U_owner.get();
  U->g = g; U->num_sms = num_sms;
  Params2& p = U->p;
  memset(&p, 0, sizeof(p));
  p.epi = epi;
  p.OFM = g.OFM; p.CB = CB; p.chb = 1; p.NPX = bNPX; p.stride2 = 0; p.deconv = 0; p.nphases = 1;
  p.WT = bWT; p.R = bR; p.P = bWT; p.PX = PX; p.PY = PY;
  p.tiles_x = (PX + bWT - 1) / bWT; p.tiles_y = (PY + bR - 1) / bR;
  p.out_x = g.out_x; p.out_y = g.out_y; p.out_word_bytes = (int)g.out_word_bytes; p.out_img_bytes = g.out_img_bytes;
  p.wstages = 1; p.wstatic = 1; p.w_bytes = w_bytes; p.ksteps = ksteps;
  p.acc_stride = CB * bNPX;
  p.acc_stages = (2 * p.acc_stride <= 512) ? 2 : 1;
  int tc = 32;
  while (tc < p.acc_stages * p.acc_stride) tc *= 2;
  p.tmem_cols = tc;
  p.idesc = make_idesc_i8(128, bNPX, 1, g.in_signed);
  p.debug = exp_int("FCB_U2_DEBUG", 0);
  p.swap = want_swap ? 1 : 0;
  p.bias_word = bias_word;
  p.idesc_swap = make_idesc_i8(128, CB * 128, g.in_signed, 1);
  p.thin_in = 1; p.S = s; p.pad = g.PAD; p.nw = nw;
  p.BWp = (s * (bWT - 1) + g.KX + 3 + 3) / 4 * 4; p.BHp = s * (bR - 1) + g.KY;
  for (int i = 0; i < 32; i++) p.toff[i] = i < nw ? (i / g.KX) * p.BWp + (i % g.KX) : 0;
  p.patch_bytes = (p.BWp * p.BHp * 4 + 127) / 128 * 128;
  // shared memory: [im2col rows x 2 sets][weights][barriers][thin-output staging][thresholds][store staging][patches x 2]
  p.nplanes = 1; p.nsets = 2; p.set_bytes = bNPX * 128;
  p.planes[0].smem_off = 0; p.planes[0].bytes = bNPX * 128;
  int off = 2 * p.set_bytes;
  p.w_off = off; off += w_bytes;
  p.bar_off = off; off += (2 * 1 + 2 * 2 + 4 + 1 + 2 * (int)U2_NPB + 1) * 8;
  off = (off + 15) & ~15;
  p.stage_off = off; off += 8 * 256;
  p.thr_off = -1; p.thr_top = thr_top;
  if (thr_bytes) { off = (off + 127) & ~127; p.thr_off = off; off += thr_bytes; }
  p.lut_off = -1;
  {
    int D = 0;
    while ((1 << D) < epi.thr_n + 1) D++;
    const int lut_bytes = CB * 128 * 256;
    if (thr_bytes && thr_top == D && epi.thr_lut && !exp_env("FCB_U2_NO_LUT")) {
      const size_t rest = (size_t)U2_NPB * p.patch_bytes + 2048;
      if ((size_t)off + lut_bytes + rest > (size_t)227 * 1024 && p.nsets == 2) {
        // single-buffer the im2col rows to make room: the layer is bound by its threshold epilogue, not by the builders
        const int saved = p.set_bytes;
        if ((size_t)off - saved + lut_bytes + rest <= (size_t)227 * 1024) {
          p.nsets = 1;
          p.w_off -= saved; p.bar_off -= saved; p.stage_off -= saved; p.thr_off -= saved;
          off -= saved;
        }
      }
      if ((size_t)off + lut_bytes + rest <= (size_t)227 * 1024) {
        off = (off + 127) & ~127;
        p.lut_off = off;
        off += lut_bytes;
      }
    }
  }
  p.stg_bytes = CB * bNPX * 128;
  const bool fast_epi = epi.act_kind == FCB_ACT_BIAS_RELU && epi.out_bits == 8 && epi.acc_bits == 8 && g.pool <= 1 && g.OFM % 128 == 0 &&
                        g.out_word_bytes == (size_t)g.OFM;
  off = (off + 1023) & ~1023;
  p.stg_off = off;
  if (fast_epi && !exp_env("FCB_U2_NO_STAGE"))
    for (int nb = 2; nb >= 1; nb--)
      if ((size_t)off + (size_t)nb * p.stg_bytes + U2_NPB * p.patch_bytes + 1024 <= (size_t)227 * 1024) { p.stg_bufs = nb; break; }
  off += p.stg_bufs * p.stg_bytes;
  if (p.swap && !p.stg_bufs) p.swap = 0;
  p.epi_alt = (p.stg_bufs == 2 && p.acc_stages == 2 && !exp_env("FCB_U2_NO_ALT")) ? 1 : 0;
  p.wl = (p.swap && p.epi_alt && bWT % 32 == 0 && !exp_env("FCB_U2_NO_WL")) ? 1 : 0;
  p.patch_off = off; off += (int)U2_NPB * p.patch_bytes;
  U->smem = (size_t)off + 1024;
  if (U->smem > 227 * 1024) return FCB_ERR_UNSUPPORTED;
  Phase2& P = p.phases[0];
  P.nkb = 1; P.px = P.py = 0;
  P.kb[0].plane = 0; P.kb[0].flags = KB_WAIT | KB_FREE; P.kb[0].a_off = 0; P.kb[0].w_k = 0; P.kb[0].d_off = 0;
  {
    const uint64_t dims[2] = {128, (uint64_t)(CB * 128)};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {128, (uint32_t)(CB * 128)};
    int rc = umma_encode_map(&U->tmW, const_cast<int8_t*>(d_w), 2, dims, strides, box);
    if (rc) return rc;
  }
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<4, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<4, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<2, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<2, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<2, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<2, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  *out = U_owner.release();
  return FCB_OK;
}

// Thin-output transposed conv plan: d_w is [9 shifts x cch chunks][16 rows][128] s8 (row = phase*4 + channel slot).
int umma2_plan_create_dthin(const Geom& g, const int8_t* d_w, const EpiParams& epi, int num_sms, Umma2Plan** out) {
  *out = nullptr;
#ifndef FCB_EXPERIMENT
  (void)g; (void)d_w; (void)epi; (void)num_sms;
  return FCB_ERR_UNSUPPORTED;  // cross-check form: experiment builds only
#else
  if (g.kind != FCB_KIND_DECONV522 || g.OFM < 3 || g.OFM > 4 || g.out_word_bytes != 4 || g.C % 128 || g.C > 256 || g.pool > 1 ||
      epi.act_kind != FCB_ACT_BIAS_RELU || epi.out_bits != 8 || epi.acc_bits != 8)
    return FCB_ERR_UNSUPPORTED;
  const int cch = g.C / 128, PX = g.IX, PY = g.IY, NPX = 256;
  int bWT = 0, bR = 0;
  double best = 1e30;
  for (int WT = 1; WT <= std::min(PX, 252); WT++) {
    const int P = WT + 2;
    if (P > g.IX + 2) continue;
    const int R = std::min(NPX / P, PY);
    if (R < 1) continue;
    const size_t plane = ((size_t)(2 * P + 2 + NPX) * 128 + 1023) / 1024 * 1024 * cch;
    if (2 * plane + 9 * cch * 2048 + 8192 > (size_t)227 * 1024) continue;
    const double tiles = (double)((PX + WT - 1) / WT) * ((PY + R - 1) / R);
    const double cost = tiles * (1.0 + 0.0005 * R) / ((double)PX * PY);
    if (cost < best * 0.999) { best = cost; bWT = WT; bR = R; }
  }
  if (!bWT) return FCB_ERR_UNSUPPORTED;
  std::unique_ptr<Umma2Plan> U_owner(new Umma2Plan());  // released to the caller on success; every early return frees it
  Umma2Plan* U = U_owner.get();
  U->g = g; U->num_sms = num_sms;
  Params2& p = U->p;
  memset(&p, 0, sizeof(p));
  p.epi = epi;
  p.OFM = g.OFM; p.CB = 1; p.chb = 1; p.NPX = NPX; p.stride2 = 0; p.deconv = 1; p.nphases = 1; p.dthin = 1; p.bias_word = -1; p.ksteps = 4;
  p.WT = bWT; p.R = bR; p.P = bWT + 2; p.PX = PX; p.PY = PY;
  p.tiles_x = (PX + bWT - 1) / bWT; p.tiles_y = (PY + bR - 1) / bR;
  p.out_x = g.out_x; p.out_y = g.out_y; p.out_word_bytes = 4; p.out_img_bytes = g.out_img_bytes;
  p.wstages = 9 * cch; p.wstatic = 1; p.w_bytes = 2048;
  p.acc_stride = 32; p.acc_stages = 2; p.tmem_cols = 64;
  p.idesc = p.idesc_dthin = make_idesc_i8(128, 16, g.in_signed, 1);
  p.debug = exp_int("FCB_U2_DEBUG", 0);
  const int rows = bR + 2;
  const size_t plane = ((size_t)(2 * p.P + 2 + NPX) * 128 + 1023) / 1024 * 1024;
  int off = 0;
  for (int cc = 0; cc < cch; cc++) {
    Plane2& pl = p.planes[cc];
    pl.smem_off = off; pl.bytes = rows * p.P * 128; pl.c0 = cc * 128; pl.dx = -1; pl.dy = -1; pl.par = 0; pl.map = 0;
    off += (int)plane;
  }
  p.nplanes = cch; p.set_bytes = off; p.nsets = 2;
  off *= 2;
  U->box_rows[0] = U->box_rows[1] = rows;
  p.w_off = off; off += p.wstages * p.w_bytes;
  p.bar_off = off; off += (2 * p.wstages + 2 * cch * 2 + 4 + 1 + 2 * (int)U2_NPB + 1) * 8;
  off = (off + 15) & ~15;
  p.stage_off = off; off += 8 * 256;
  p.thr_off = -1; p.lut_off = -1;
  U->smem = (size_t)off + 1024;
  if (U->smem > 227 * 1024) return FCB_ERR_UNSUPPORTED;
  // K-blocks: (channel chunk, shift); shift s = (offy + 1) * 3 + (offx + 1)
  Phase2& P = p.phases[0];
  P.nkb = 0; P.px = P.py = 0;
  for (int cc = 0; cc < cch; cc++)
    for (int sft = 0; sft < 9; sft++) {
      KB2& kb = P.kb[P.nkb++];
      kb.plane = (uint16_t)cc;
      kb.flags = (sft == 0 ? KB_WAIT : 0) | (sft == 8 ? KB_FREE : 0);
      kb.a_off = (sft / 3) * p.P + (sft % 3);
      kb.w_k = (cc * 9 + sft) * 16;  // first weight row of the block
      kb.d_off = (uint32_t)(p.planes[cc].smem_off + kb.a_off * 128) >> 4;
    }
  {
    const uint64_t dims[2] = {128, (uint64_t)(9 * cch * 16)};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {128, 16};
    int rc = umma_encode_map(&U->tmW, const_cast<int8_t*>(d_w), 2, dims, strides, box);
    if (rc) return rc;
  }
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  *out = U_owner.release();
  return FCB_OK;
#endif
}

// Thin-output transposed conv, col2im form: d_w is [cch][128 rows (tap*OFM + ch, zero beyond 25*OFM)][128] s8.
int umma2_plan_create_dcol(const Geom& g, const int8_t* d_w, const EpiParams& epi, int num_sms, Umma2Plan** out) {
  *out = nullptr;
  if (g.kind != FCB_KIND_DECONV522 || g.OFM < 3 || g.OFM > 4 || g.out_word_bytes != 4 || g.C % 128 || g.C > 256 || g.pool > 1 ||
      epi.act_kind != FCB_ACT_BIAS_RELU || epi.out_bits != 8 || epi.acc_bits != 8)
    return FCB_ERR_UNSUPPORTED;
  const int cch = g.C / 128, PX = g.IX, PY = g.IY, NPX = 256;
  int bWT = 0, bR = 0;
  double best = 1e30;
  for (int WT = 1; WT <= std::min(PX, 254); WT++) {
    const int P = WT + 2;
    const int R = std::min(NPX / P - 2, PY);
    if (R < 1) continue;
    const double tiles = (double)((PX + WT - 1) / WT) * ((PY + R - 1) / R);
    const double cost = tiles / ((double)PX * PY);  // every tile costs the same (256 columns of MMA, TMEM read and col2im)
    if (cost < best * 0.999) { best = cost; bWT = WT; bR = R; }
  }
  if (!bWT) return FCB_ERR_UNSUPPORTED;
  std::unique_ptr<Umma2Plan> U_owner(new Umma2Plan());  // released to the caller on success; every early return frees it
  Umma2Plan* U = U_owner.get();
  U->g = g; U->num_sms = num_sms;
  Params2& p = U->p;
  memset(&p, 0, sizeof(p));
  p.epi = epi;
  p.OFM = g.OFM; p.CB = 1; p.chb = 1; p.NPX = NPX; p.stride2 = 0; p.deconv = 1; p.nphases = 1; p.dthin = 2; p.bias_word = -1; p.ksteps = 4;
  p.WT = bWT; p.R = bR; p.P = bWT + 2; p.PX = PX; p.PY = PY;
  p.tiles_x = (PX + bWT - 1) / bWT; p.tiles_y = (PY + bR - 1) / bR;
  p.out_x = g.out_x; p.out_y = g.out_y; p.out_word_bytes = 4; p.out_img_bytes = g.out_img_bytes;
  p.wstages = cch; p.wstatic = 1; p.w_bytes = 128 * 128;
  p.acc_stride = 2 * 112; p.acc_stages = 2; p.tmem_cols = 512;  // per stage: two 128-pixel blocks x 112 columns
  p.idesc = p.idesc_dthin = make_idesc_i8(128, 112, g.in_signed, 1);  // A = plane rows (pixels), B = the (tap word, channel) weight rows
  p.debug = exp_int("FCB_U2_DEBUG", 0);
  p.epi_alt = 1;  // one epilogue group per accumulator stage
  p.s_pitch = DCOL_PIX;  // pixel-major byte tile: 112 bytes per pixel (25 tap words + pad)
  const int rows = bR + 2;
  int off = 0;
  for (int cc = 0; cc < cch; cc++) {
    Plane2& pl = p.planes[cc];
    pl.smem_off = off; pl.bytes = rows * p.P * 128; pl.c0 = cc * 128; pl.dx = -1; pl.dy = -1; pl.par = 0; pl.map = 0;
    off += NPX * 128;
  }
  // the whole tile is one K-block, so nothing inside a tile hides the ~2.5 k clocks a 32 KB plane takes to arrive from HBM:
  // the planes are prefetched three tiles ahead
  p.nplanes = cch; p.set_bytes = off;
  p.nsets = 4;  // measured on L7: 2 sets 201.5 k, 3 sets 209 k, 4 sets 218 k, 5 sets 212 k img/s
  p.nsets = std::max(1, std::min(5, exp_int("FCB_U2_NSETS", p.nsets)));
  while (p.nsets > 1 && (size_t)p.nsets * off + (size_t)cch * 16384 + 2 * (NPX * DCOL_PIX) + 8192 > (size_t)227 * 1024) p.nsets--;
  off *= p.nsets;
  U->box_rows[0] = U->box_rows[1] = rows;
  p.w_off = off; off += p.wstages * p.w_bytes;
  p.bar_off = off; off += (2 * p.wstages + 2 * cch * p.nsets + 4 + 1 + 2 * (int)U2_NPB + 1) * 8;
  off = (off + 15) & ~15;
  p.stage_off = off; off += 8 * 256;
  p.thr_off = -1; p.lut_off = -1;
  off = (off + 127) & ~127;
  p.stg_off = off; p.stg_bytes = NPX * DCOL_PIX; p.stg_bufs = 0;  // (stg_bufs stays 0: not the TMA-store path)
  off += 2 * p.stg_bytes;
  U->smem = (size_t)off + 1024;
  if (U->smem > 227 * 1024) return FCB_ERR_UNSUPPORTED;
  Phase2& P = p.phases[0];
  P.nkb = 0; P.px = P.py = 0;
  for (int cc = 0; cc < cch; cc++) {
    KB2& kb = P.kb[P.nkb++];
    kb.plane = (uint16_t)cc; kb.flags = KB_WAIT | KB_FREE; kb.a_off = 0; kb.w_k = cc * 128;  // weight rows cc*128 .. +127
    kb.d_off = (uint32_t)p.planes[cc].smem_off >> 4;
  }
  {
    const uint64_t dims[2] = {128, (uint64_t)(cch * 128)};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {128, 128};
    int rc = umma_encode_map(&U->tmW, const_cast<int8_t*>(d_w), 2, dims, strides, box);
    if (rc) return rc;
  }
#ifdef FCB_EXPERIMENT
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#endif
  FCB_CUDA_OK(cudaFuncSetAttribute(umma2_conv_kernel<1, 2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  *out = U_owner.release();
  return FCB_OK;
}

void umma2_plan_destroy(Umma2Plan* U) { delete U; }

static const char* umma2_describe_base(const Umma2Plan* U, char* buf, size_t n);
// Pooled 8-bit threshold layers with a monotone compare and shared-memory tables run the instantiation that contains nothing else
// (EPI >= 1).  Returns -1 for other layers, else the number of EXTRA epilogue warps (0, 4, 8).
// The lock-step search is latency-bound (issue slots ~40 % busy with 8 epilogue warps): tiles with enough (row pair, 16-column)
// units get the group count that minimises units per group.  Measured (img/s with 8 / 12 / 16 epilogue warps): config 5b stage 1
// (8 units) 21.5 k / 24.2 k / 25.8 k, stage 2 (6 units) 60.1 k / 69.3 k / 67.5 k, config 4 (4 units) 583 k / 532 k / 560 k.
static int thrp_extra_warps(const Params2& p) {
  const EpiParams& e = p.epi;
  const bool thrp = e.act_kind == FCB_ACT_THRESHOLDS && e.pool == 2 && e.out_bits == 8 && p.thr_off >= 0 && !p.deconv &&
                    (e.cmp == FCB_CMP_LESS || e.cmp == FCB_CMP_LESS_EQUAL) && e.act_val >= 0 && e.act_val + e.num_th < 256 &&
                    (e.acc_signed || e.acc_bits < 32) && !exp_env("FCB_U2_NO_THRP");
  if (!thrp) return -1;
  if (exp_env("FCB_U2_XEPI")) { const int xe = exp_int("FCB_U2_XEPI", 0); return xe >= 8 ? 8 : xe >= 4 ? 4 : 0; }  // experiment switch
  const int nunits = (p.R >> 1) * ((p.WT + 15) >> 4);
  return nunits < 6 ? 0 : (nunits + 3) / 4 < (nunits + 2) / 3 ? 8 : 4;
}

const char* umma2_describe(const Umma2Plan* U, char* buf, size_t n) {
  const char* d = umma2_describe_base(U, buf, n);
  const int xw = thrp_extra_warps(U->p);
  const size_t len = strlen(buf);
  if (xw >= 0 && len + 24 < n) snprintf(buf + len, n - len, " epi-warps=%d", 8 + xw);
  return d;
}

static const char* umma2_describe_base(const Umma2Plan* U, char* buf, size_t n) {
  const Params2& p = U->p;
  if (p.dthin == 2) {
    snprintf(buf, n, "thin-output deconv: GEMM over (tap, channel) rows + shared-memory col2im, WT=%d R=%d (+1 halo ring, %d of 256 columns) planes=%dx%d smem=%zu tiles=%dx%d",
             p.WT, p.R, (p.WT + 2) * (p.R + 2), p.nplanes, p.nsets, U->smem, p.tiles_x, p.tiles_y);
    return buf;
  }
  if (p.dthin) {
    snprintf(buf, n, "thin-output deconv: pixels on M, 9 shift blocks x N=16 (4 phases x 4 ch) WT=%d R=%d P=%d planes=%dx%d smem=%zu tiles=%dx%d",
             p.WT, p.R, p.P, p.nplanes, p.nsets, U->smem, p.tiles_x, p.tiles_y);
    return buf;
  }
  if (p.thin_in) {
    snprintf(buf, n, "smem-im2col (thin input, K=%d B%s, %d MMA/tile, weights resident) WT=%d R=%d NPX=%d CB=%d%s patch=%dx%d smem=%zu tiles=%dx%d",
             p.nw * 4, p.bias_word >= 0 ? " + bias row" : "", p.swap ? p.ksteps * (p.NPX / 128) : p.ksteps * p.CB, p.WT, p.R, p.NPX, p.CB,
             p.lut_off >= 0 ? " thr@smem + bucket LUT" : p.thr_off >= 0 ? " thr@smem" : (p.wl ? " pixel-major warp-local tma-store" : p.swap ? " pixel-major tma-store" : p.stg_bufs == 2 ? " tma-store x2" : p.stg_bufs == 1 ? " tma-store x1" : ""), p.BWp, p.BHp, U->smem,
             p.tiles_x, p.tiles_y);
    return buf;
  }
  snprintf(buf, n, "resident-planes WT=%d R=%d P=%d NPX=%d N=%d CB=%d chb=%d%s%s planes=%dx%d wstages=%d acc_stages=%d smem=%zu tiles=%dx%d", p.WT,
           p.R, p.P, p.NPX, p.n_mma, p.CB, p.chb, p.lut_off >= 0 ? " thr@smem + bucket LUT" : p.thr_off >= 0 ? " thr-top@smem" : (p.stg_bufs == 2 ? " tma-store x2" : p.stg_bufs == 1 ? " tma-store x1" : p.epi4 ? " epi-warps=4" : ""), U->cluster_ok ? " weights-multicast x2" : "", p.nplanes, p.nsets, p.wstages, p.acc_stages, U->smem,
           p.tiles_x, p.tiles_y);
  return buf;
}

static int umma2_encode_maps(const Umma2Plan* U, MapSet* ms, const void* d_in, void* d_out, int n_images) {
  const Geom& g = U->g;
  const Params2& p = U->p;
  const uint64_t C = g.C, X = g.IX, Y = g.IY;
  for (int m = 0; m < 2; m++) {
    int rc;
    if (p.thin_in) {  // raw image of 4-byte pixels; the patch of a tile is one box, borders zero-filled
      const uint64_t dims[3] = {X, Y, (uint64_t)n_images};
      const uint64_t strides[2] = {X * 4, X * Y * 4};
      const uint32_t box[3] = {(uint32_t)p.BWp, (uint32_t)p.BHp, 1};
      rc = umma_encode_map_ex(&ms->tmA[m], const_cast<void*>(d_in), 4, 0, 3, dims, strides, box);
    } else if (p.stride2) {
      const uint64_t dims[5] = {2 * C, X / 2, 2, Y / 2, (uint64_t)n_images};
      const uint64_t strides[4] = {2 * C, X * C, 2 * X * C, X * Y * C};
      const uint32_t box[5] = {128, (uint32_t)p.P, 1, U->box_rows[m], 1};
      rc = umma_encode_map(&ms->tmA[m], const_cast<void*>(d_in), 5, dims, strides, box);
    } else {
      const uint64_t dims[4] = {C, X, Y, (uint64_t)n_images};
      const uint64_t strides[3] = {C, X * C, X * Y * C};
      const uint32_t box[4] = {128, (uint32_t)p.P, U->box_rows[m], 1};
      rc = umma_encode_map(&ms->tmA[m], const_cast<void*>(d_in), 4, dims, strides, box);
    }
    if (rc) return rc;
  }
  ms->tmO = ms->tmA[0];
  if (p.stg_bufs > 0) {
    const uint64_t F = g.OFM, OX = g.out_x, OY = g.out_y;
    int rc;
    if (p.deconv) {  // [n][oy/2][oy%2][ox/2][(ox%2)*OFM + ch]: a phase's tile row is a dense box
      const uint64_t dims[5] = {2 * F, OX / 2, 2, OY / 2, (uint64_t)n_images};
      const uint64_t strides[4] = {2 * F, OX * F, 2 * OX * F, OX * OY * F};
      const uint32_t box[5] = {128, (uint32_t)p.WT, 1, 1, 1};
      rc = umma_encode_map_ex(&ms->tmO, d_out, 1, 0, 5, dims, strides, box);
    } else {
      const uint64_t dims[4] = {F, OX, OY, (uint64_t)n_images};
      const uint64_t strides[3] = {F, OX * F, OX * OY * F};
      const uint32_t box[4] = {128, (uint32_t)(p.wl ? 32 : p.WT), 1, 1};
      rc = umma_encode_map_ex(&ms->tmO, d_out, 1, p.swap ? 128 : 0, 4, dims, strides, box);
    }
    if (rc) return rc;
  }
  ms->d_in = d_in; ms->d_out = d_out; ms->n_images = n_images;
  return FCB_OK;
}

int umma2_run(Umma2Plan* U, const void* d_in, void* d_out, int n_images, cudaStream_t st) {
  const Geom& g = U->g;
  Params2 p = U->p;
  p.out = (uint8_t*)d_out;
  p.n_images = n_images;
  const MapSet* ms = nullptr;
  for (const MapSet& c : U->maps)
    if (c.n_images == n_images && c.d_in == d_in && c.d_out == d_out) { ms = &c; break; }
  if (!ms) {
    MapSet* slot = &U->maps[U->map_next];
    slot->n_images = 0;  // invalid while it is being rewritten
    int rc = umma2_encode_maps(U, slot, d_in, d_out, n_images);
    if (rc) return rc;
    U->map_next = (U->map_next + 1) & 3;
    ms = slot;
  }
  const CUtensorMap* tmA = ms->tmA;
  const CUtensorMap& tmO = ms->tmO;
  // sub-byte / padded output words are merged or partially written: start from zeroed words
  if (g.out_word_bytes * 8 != (size_t)g.OFM * g.out_bits && !p.dthin)  // (the thin-output deconv epilogue writes whole words)
    FCB_CUDA_OK(cudaMemsetAsync(d_out, 0, g.out_img_bytes * n_images, st));
  const long long total = (long long)p.tiles_x * p.tiles_y * n_images;
  const int grid = (int)std::min<long long>(total, U->num_sms / p.chb) * p.chb;
  unsigned long long* d_prof = nullptr;
#ifdef FCB_U2_PROF
  if (exp_env("FCB_U2_PROF")) {
    FCB_CUDA_OK(cudaMalloc(&d_prof, (size_t)grid * 24 * 8));
    FCB_CUDA_OK(cudaMemsetAsync(d_prof, 0, (size_t)grid * 24 * 8, st));
    p.prof = d_prof;
  }
#endif
  // thin-input: 4 builder warps beside the light bias/ReLU epilogue (128 registers per thread suffice); 2 beside the threshold
  // epilogue, whose lock-step searches need ~170 registers to stay out of local memory
  const int xepi = thrp_extra_warps(p);
  const bool thrp = xepi >= 0;
  if (p.thin_in && thrp && xepi == 8) umma2_conv_kernel<2, 0, 3><<<grid, 320 + 32 * 2 + 256, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.thin_in && thrp && xepi == 4) umma2_conv_kernel<2, 0, 2><<<grid, 320 + 32 * 2 + 128, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.thin_in && thrp) umma2_conv_kernel<2, 0, 1><<<grid, 320 + 32 * 2, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.thin_in && p.epi.act_kind == FCB_ACT_THRESHOLDS) umma2_conv_kernel<2, 0, 0><<<grid, 320 + 32 * 2, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.thin_in && p.swap && p.wl && p.epi_alt && p.NPX == 256 && !exp_env("FCB_U2_NO_SWPX"))  // pixel-major bias+ReLU, warp-local stores: 16 epilogue warps
    umma2_conv_kernel<4, 0, 4><<<grid, 320 + 32 * 4 + 256, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.thin_in) umma2_conv_kernel<4, 0, 0><<<grid, 320 + 32 * 4, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.dthin == 2 && !exp_env("FCB_U2_NO_DCX")) umma2_conv_kernel<1, 2, 5><<<grid, 320 + 32 + 256, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
#ifdef FCB_EXPERIMENT
  else if (p.dthin == 2) umma2_conv_kernel<1, 2, 0><<<grid, 320 + 32, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
  else if (p.dthin) umma2_conv_kernel<1, 1, 0><<<grid, 320 + 32, U->smem, st>>>(tmA[0], tmA[1], U->tmW, tmO, p);
#endif
  else if (!thrp && p.stg_bufs == 2 && p.epi_alt && !p.swap && !exp_env("FCB_U2_NO_STX"))  // staged bias+ReLU, two tiles: 16 epilogue warps
    FCB_CUDA_OK(launch_resident<6>(U, grid, 320 + 32 + 256, st, tmA[0], tmA[1], tmO, p));
  else if (thrp && xepi == 8) FCB_CUDA_OK(launch_resident<3>(U, grid, 320 + 32 + 256, st, tmA[0], tmA[1], tmO, p));
  else if (thrp && xepi == 4) FCB_CUDA_OK(launch_resident<2>(U, grid, 320 + 32 + 128, st, tmA[0], tmA[1], tmO, p));
  else if (thrp) FCB_CUDA_OK(launch_resident<1>(U, grid, 320 + 32, st, tmA[0], tmA[1], tmO, p));
  else FCB_CUDA_OK(launch_resident<0>(U, grid, 320 + 32, st, tmA[0], tmA[1], tmO, p));
  FCB_CUDA_OK(cudaGetLastError());
  if (d_prof) {  // debugging aid: average clocks per tile and segment over the CTAs
    FCB_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<unsigned long long> h((size_t)grid * 24);
    FCB_CUDA_OK(cudaMemcpy(h.data(), d_prof, h.size() * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_prof);
    const double tiles_per_cta = (double)total / grid;
    const char* names[3] = {"builder", "mma", "epilogue"};
    for (int role = 0; role < 3; role++) {
      fprintf(stderr, "[u2 prof] %-8s clk/tile:", names[role]);
      for (int sgm = 0; sgm < 8; sgm++) {
        double sum = 0;
        for (int c = 0; c < grid; c++) sum += (double)h[((size_t)c * 3 + role) * 8 + sgm];
        fprintf(stderr, " %7.0f", sum / grid / tiles_per_cta);
      }
      fprintf(stderr, "\n");
    }
  }
  return FCB_OK;
}

}  // namespace fcb
