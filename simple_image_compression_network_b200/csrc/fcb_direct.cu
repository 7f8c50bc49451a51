// fcb_direct.cu -- direct (CUDA-core) convolution engines: "imad" and "xnor_popc".
//
// imad      : the universal integer path (any lane width <= 16 bits, +-1 / xnor weights, any activation).
//             north_star: "IMAD for wider types".  Replaces, in one kernel,
//             FMPadding_nonsquare (streamtools.h:361-406), the zero-insertion of deconv522
//             (conv_nonsquare_top.cpp:109-156), ConvolutionInputGenerator_NonSquare
//             (slidingwindow.h:1254-1353), Matrix_Vector_Activate_Batch (mvau.hpp:87-179) and the
//             activation / bias-ReLU / pool stages.
// xnor_popc : BinaryWeights + Recast<XnorMul> (weights.hpp:66-98, interpret.hpp:57-73) on bit-packed
//             lanes: acc = sum over 32-bit words of popc(~(a ^ w)).
//
// Work decomposition (both): one CTA = 16x8 pre-pool output pixels x 64 output channels of one
// image; a warp owns a 4x4 pixel block, a lane owns channels {l, l+32}; the input patch of the tile
// is staged once in shared memory (zero padding / zero insertion resolved while staging), weights
// are read through L1 as [k][channel] rows so a warp's 32 lanes load 32 consecutive values.
#include <algorithm>
#include <vector>

#include "fcb_epilogue.cuh"

namespace fcb {

constexpr int TX = 16, TY = 8, PXB = 4;  // tile and per-warp pixel block
constexpr int CH_PER_CTA = 64;

// Maps a coordinate of the virtual padded frame to the input image; false = structural zero.
__device__ __forceinline__ bool map_coord(int v, int pad, int deconv, int extent, int* src) {
  if (!deconv) {
    const int s = v - pad;
    *src = s;
    return s >= 0 && s < extent;
  }
  const int s = v - 2;  // deconv522: Z(2i,2j) = a(i,j), frame padded by 2
  *src = s >> 1;
  return s >= 0 && !(s & 1) && (s >> 1) < extent;
}

// activation, pool, store of one warp's 4x4 pixel block x 2 channels per lane (shared by both direct kernels)
__device__ __forceinline__ void direct_epilogue(const DirectParams& p, const int32_t (&acc)[PXB * PXB][2], int img, int ch0, int ox0, int oy0, int wx,
                                                int wy, int lane) {
  // ---- activation, pool, store -----------------------------------------------------------
  const int pk = p.epi.pool >= 2 ? p.epi.pool : 1;
  uint8_t* out = p.out + (size_t)img * p.out_img_bytes;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int ch = ch0 + 32 * j;
    const bool chv = ch < p.OFM;
    uint32_t val[PXB * PXB];
#pragma unroll
    for (int i = 0; i < PXB * PXB; i++) val[i] = chv ? activate(p.epi, ch, acc[i][j]) : 0u;
    if (p.epi.out_bits == 1 && pk == 1) {
      // 1-bit lanes: one ballot per pixel; lane i keeps pixel i of the 4x4 block, then one store instruction for all 16
      uint32_t mine = 0;
#pragma unroll
      for (int i = 0; i < PXB * PXB; i++) {
        const uint32_t bits = __ballot_sync(0xffffffffu, chv && (val[i] & 1u));
        if (lane == i) mine = bits;
      }
      const int oy = oy0 + wy + (lane >> 2), ox = ox0 + wx + (lane & 3), c0 = ch - lane;  // c0: first channel of this warp
      if (lane < PXB * PXB && oy < p.OY && ox < p.OX && c0 < p.OFM) {
        uint8_t* dst = out + ((size_t)oy * p.out_x + ox) * p.out_word_bytes + (c0 >> 3);
        if (c0 + 32 <= p.OFM && p.out_word_bytes >= 4) *reinterpret_cast<uint32_t*>(dst) = mine;
        else
          for (int b = 0; b < 4; b++)
            if (c0 + 8 * b < p.OFM) dst[b] = (uint8_t)(mine >> (8 * b));
      }
      continue;
    }
    for (int by = 0; by < PXB; by += pk)
      for (int bx = 0; bx < PXB; bx += pk) {
        const int oy = oy0 + wy + by, ox = ox0 + wx + bx;  // pre-pool pixel (warp-uniform)
        if (oy >= p.OY || ox >= p.OX) continue;
        uint32_t m = 0;
        for (int dy = 0; dy < pk; dy++)
          for (int dx = 0; dx < pk; dx++) m = max(m, val[(by + dy) * PXB + bx + dx]);
        uint8_t* word = out + ((size_t)(oy / pk) * p.out_x + (ox / pk)) * p.out_word_bytes;
        store_lane(word, ch, chv, m, p.epi.out_bits);
      }
  }
}

template <int ENGINE>
__global__ void __launch_bounds__(256) direct_conv_kernel(const DirectParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int img = blockIdx.z;
  const int ch0 = blockIdx.y * CH_PER_CTA + lane;
  const int ox0 = tx * TX, oy0 = ty * TY;
  const int wx = (warp & 3) * PXB, wy = (warp >> 2) * PXB;  // this warp's 4x4 pixel block in the tile
  const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;

  int32_t acc[PXB * PXB][2];
#pragma unroll
  for (int i = 0; i < PXB * PXB; i++) acc[i][0] = acc[i][1] = 0;

  const int vx0 = ox0 * p.SXe, vy0 = oy0 * p.SYe;  // tile origin in the padded frame
  // channel units: imad = lanes, xnor = 32-bit words of packed lanes
  const int CU = (ENGINE == ENG_XNOR) ? (p.C >> 5) : p.C;

  for (int c0 = 0; c0 < CU; c0 += p.CC) {
    const int cc = min(p.CC, CU - c0);
    // ---- stage the patch ------------------------------------------------------------
    const int total = p.patch_h * p.patch_w * cc;
    for (int idx = tid; idx < total; idx += blockDim.x) {
      const int c = idx % cc, pix = idx / cc;
      const int px = pix % p.patch_w, py = pix / p.patch_w;
      int sx, sy;
      const bool okx = map_coord(vx0 + px, p.PAD, p.deconv, p.IX, &sx);
      const bool oky = map_coord(vy0 + py, p.PADY, p.deconv, p.IY, &sy);  // (FMPadding_nonsquare may pad left / up differently)
      if (ENGINE == ENG_XNOR) {
        uint32_t v = 0;
        if (okx && oky) v = reinterpret_cast<const uint32_t*>(in + ((size_t)sy * p.IX + sx) * p.in_word_bytes)[c0 + c];
        reinterpret_cast<uint32_t*>(smem_raw)[idx] = v;
      } else {
        int32_t v = 0;
        if (okx && oky) v = load_lane_any(in + ((size_t)sy * p.IX + sx) * p.in_word_bytes, c0 + c, p.in_bits, p.in_signed);
        reinterpret_cast<int32_t*>(smem_raw)[idx] = v;
      }
    }
    __syncthreads();
    // ---- multiply-accumulate ----------------------------------------------------------
    if (ENGINE == ENG_XNOR && (cc & 1) == 0) {
      // channel words in pairs: one 8-byte patch load feeds two popcounts per output, the two counts and the accumulator meet
      // in one 3-input add, and the weights of the next (tap, pair) are fetched while this one is counted
      const uint32_t* wt = reinterpret_cast<const uint32_t*>(p.wt);
      const uint2* patch2 = reinterpret_cast<const uint2*>(smem_raw);
      const int cc2 = cc >> 1, nit = p.KY * p.KX * cc2;
      int ky = 0, kx = 0, c2 = 0;
      auto wrow_of = [&](int ky_, int kx_, int c2_) { return (size_t)((ky_ * p.KX + kx_) * CU + c0 + 2 * c2_) * p.OFMp + ch0; };
      uint32_t wn[4];
      {
        const size_t r = wrow_of(0, 0, 0);
        wn[0] = __ldg(wt + r); wn[1] = __ldg(wt + r + 32); wn[2] = __ldg(wt + r + p.OFMp); wn[3] = __ldg(wt + r + p.OFMp + 32);
      }
      for (int it = 0; it < nit; it++) {
        const uint32_t w00 = wn[0], w01 = wn[1], w10 = wn[2], w11 = wn[3];
        const int poff = (ky * p.DY * p.patch_w + kx * p.DX) * cc2 + c2;
        if (++c2 == cc2) { c2 = 0; if (++kx == p.KX) { kx = 0; ++ky; } }
        if (it + 1 < nit) {
          const size_t r = wrow_of(ky, kx, c2);
          wn[0] = __ldg(wt + r); wn[1] = __ldg(wt + r + 32); wn[2] = __ldg(wt + r + p.OFMp); wn[3] = __ldg(wt + r + p.OFMp + 32);
        }
#pragma unroll
        for (int i = 0; i < PXB * PXB; i++) {
          const int ly = wy + (i >> 2), lx = wx + (i & 3);
          const uint2 a = patch2[(ly * p.SYe * p.patch_w + lx * p.SXe) * cc2 + poff];
          acc[i][0] += __popc(~(a.x ^ w00)) + __popc(~(a.y ^ w10));
          acc[i][1] += __popc(~(a.x ^ w01)) + __popc(~(a.y ^ w11));
        }
      }
    } else
    for (int ky = 0; ky < p.KY; ky++)
      for (int kx = 0; kx < p.KX; kx++) {
        const int kbase = (ky * p.KX + kx) * CU + c0;
        for (int c = 0; c < cc; c++) {
          const size_t wrow = (size_t)(kbase + c) * p.OFMp + ch0;
          if (ENGINE == ENG_XNOR) {
            const uint32_t* wt = reinterpret_cast<const uint32_t*>(p.wt);
            const uint32_t w0 = __ldg(wt + wrow), w1 = __ldg(wt + wrow + 32);
            const uint32_t* patch = reinterpret_cast<const uint32_t*>(smem_raw);
#pragma unroll
            for (int i = 0; i < PXB * PXB; i++) {
              const int ly = wy + (i >> 2), lx = wx + (i & 3);
              const uint32_t a = patch[((ly * p.SYe + ky * p.DY) * p.patch_w + (lx * p.SXe + kx * p.DX)) * cc + c];
              acc[i][0] += __popc(~(a ^ w0));
              acc[i][1] += __popc(~(a ^ w1));
            }
          } else {
            const int16_t* wt = reinterpret_cast<const int16_t*>(p.wt);
            const int32_t w0 = __ldg(wt + wrow), w1 = __ldg(wt + wrow + 32);
            const int32_t* patch = reinterpret_cast<const int32_t*>(smem_raw);
#pragma unroll
            for (int i = 0; i < PXB * PXB; i++) {
              const int ly = wy + (i >> 2), lx = wx + (i & 3);
              const int32_t a = patch[((ly * p.SYe + ky * p.DY) * p.patch_w + (lx * p.SXe + kx * p.DX)) * cc + c];
              if (p.mul_kind == FCB_W_BINARY_XNOR) {
                acc[i][0] += (a == w0);
                acc[i][1] += (a == w1);
              } else {
                acc[i][0] += a * w0;
                acc[i][1] += a * w1;
              }
            }
          }
        }
      }
    __syncthreads();
  }
  direct_epilogue(p, acc, img, ch0, ox0, oy0, wx, wy, lane);
}

// ------------------------------------------------------------------------------------------------------------------------
// imad_conv_kernel -- the IMAD engine proper (FixedPoint / +-1 weights; north_star: "IMAD for wider types").
// Same decomposition as above (CTA = 16x8 pixels x 64 channels, warp = 4x4 pixels, lane = channels {l, l+32}), with the inner loop
// built around the shared-memory pipe instead of L1:
//   * the weights of the channel chunk are staged in shared memory as packed pairs, word [tap][c/4][lane][c%4] = w(ch l) | w(ch l+32) << 16:
//     one conflict-free LDS.128 per (tap, 4 channels) instead of eight L1 loads whose latency nothing hid;
//   * the patch is stored with a channel pitch that is a multiple of 4, so one LDS.128 brings 4 channels of a pixel: per 4 channels
//     a lane issues 16 + 1 shared loads for 128 IMADs (16 + 2 per 32 before);
//   * deconv522: taps are walked by parity class and only the pixels whose (row + ky, column + kx) parities hit a non-zero sample
//     of the zero-inserted frame are unrolled: 4x fewer IMADs, the structural zeros are never multiplied.
template <bool DECONV, int PKY, int PKX>
__device__ __forceinline__ void imad_taps(const DirectParams& p, const int32_t* __restrict__ patch, const uint4* __restrict__ wsm, int cc4, int lane,
                                          int wx, int wy, int32_t (&acc)[PXB * PXB][2]) {
  const int rowstep = p.SYe * p.patch_w * cc4, colstep = p.SXe * cc4, ngrp = cc4 >> 2;
  for (int ky = DECONV ? PKY : 0; ky < p.KY; ky += DECONV ? 2 : 1)
    for (int kx = DECONV ? PKX : 0; kx < p.KX; kx += DECONV ? 2 : 1) {
      const uint4* wrow = wsm + (size_t)(ky * p.KX + kx) * ngrp * 32 + lane;
      const int32_t* prow = patch + ((size_t)(wy * p.SYe + ky * p.DY) * p.patch_w + wx * p.SXe + kx * p.DX) * cc4;
      for (int g = 0; g < ngrp; g++) {
        const uint4 wp = wrow[g * 32];  // 4 channels x {ch l, ch l+32}: one conflict-free LDS.128
        const int32_t w00 = (int16_t)(wp.x & 0xFFFFu), w01 = (int16_t)(wp.y & 0xFFFFu), w02 = (int16_t)(wp.z & 0xFFFFu), w03 = (int16_t)(wp.w & 0xFFFFu);
        const int32_t w10 = (int32_t)wp.x >> 16, w11 = (int32_t)wp.y >> 16, w12 = (int32_t)wp.z >> 16, w13 = (int32_t)wp.w >> 16;
#pragma unroll
        for (int i = 0; i < PXB * PXB; i++) {
          const int ly = i >> 2, lx = i & 3;
          if (DECONV && (((ly ^ PKY) & 1) || ((lx ^ PKX) & 1))) continue;  // compile-time: structural zero of the inserted frame
          const int4 a = *reinterpret_cast<const int4*>(prow + ly * rowstep + lx * colstep + g * 4);
          // one IMAD per MAC, chained on the accumulator (a sum tree costs an extra add per four)
          int32_t s0 = acc[i][0], s1 = acc[i][1];
          s0 = a.x * w00 + s0; s1 = a.x * w10 + s1;
          s0 = a.y * w01 + s0; s1 = a.y * w11 + s1;
          s0 = a.z * w02 + s0; s1 = a.z * w12 + s1;
          s0 = a.w * w03 + s0; s1 = a.w * w13 + s1;
          acc[i][0] = s0; acc[i][1] = s1;
        }
      }
    }
}

__global__ void __launch_bounds__(256, 2) imad_conv_kernel(const DirectParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int img = blockIdx.z;
  const int chb = blockIdx.y * CH_PER_CTA, ch0 = chb + lane;
  const int ox0 = tx * TX, oy0 = ty * TY;
  const int wx = (warp & 3) * PXB, wy = (warp >> 2) * PXB;
  const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;
  int32_t acc[PXB * PXB][2];
#pragma unroll
  for (int i = 0; i < PXB * PXB; i++) acc[i][0] = acc[i][1] = 0;
  const int vx0 = ox0 * p.SXe, vy0 = oy0 * p.SYe;
  int32_t* patch = reinterpret_cast<int32_t*>(smem_raw);
  const int16_t* wt = reinterpret_cast<const int16_t*>(p.wt);
  for (int c0 = 0; c0 < p.C; c0 += p.CC) {
    const int cc = min(p.CC, p.C - c0), cc4 = (cc + 3) & ~3;
    uint32_t* wsm = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)p.patch_h * p.patch_w * cc4;
    const uint4* wsm4 = reinterpret_cast<const uint4*>(wsm);
    // ---- stage the patch: zero padding / zero insertion resolved here; channels cc..cc4 are zeros
    const int total = p.patch_h * p.patch_w * cc4;
    for (int idx = tid; idx < total; idx += blockDim.x) {
      const int c = idx % cc4, pix = idx / cc4;
      const int px = pix % p.patch_w, py = pix / p.patch_w;
      int sx, sy;
      const bool okx = map_coord(vx0 + px, p.PAD, p.deconv, p.IX, &sx);
      const bool oky = map_coord(vy0 + py, p.PADY, p.deconv, p.IY, &sy);
      int32_t v = 0;
      if (c < cc && okx && oky) v = load_lane_any(in + ((size_t)sy * p.IX + sx) * p.in_word_bytes, c0 + c, p.in_bits, p.in_signed);
      patch[idx] = v;
    }
    // ---- stage the weights of this channel chunk as packed pairs, word [tap][c / 4][lane][c % 4]
    const int wtotal = p.KX * p.KY * cc4 * 32;
    for (int idx = tid; idx < wtotal; idx += blockDim.x) {
      const int l = (idx >> 2) & 31, grp = idx >> 7, c = (grp % (cc4 >> 2)) * 4 + (idx & 3), tap = grp / (cc4 >> 2);
      uint32_t wp = 0;
      if (c < cc) {
        const size_t row = (size_t)(tap * p.C + c0 + c) * p.OFMp + chb + l;
        wp = ((uint32_t)(uint16_t)__ldg(wt + row)) | ((uint32_t)(uint16_t)__ldg(wt + row + 32) << 16);
      }
      wsm[idx] = wp;
    }
    __syncthreads();
    if (p.deconv) {
      imad_taps<true, 0, 0>(p, patch, wsm4, cc4, lane, wx, wy, acc);
      imad_taps<true, 0, 1>(p, patch, wsm4, cc4, lane, wx, wy, acc);
      imad_taps<true, 1, 0>(p, patch, wsm4, cc4, lane, wx, wy, acc);
      imad_taps<true, 1, 1>(p, patch, wsm4, cc4, lane, wx, wy, acc);
    } else {
      imad_taps<false, 0, 0>(p, patch, wsm4, cc4, lane, wx, wy, acc);
    }
    __syncthreads();
  }
  direct_epilogue(p, acc, img, ch0, ox0, oy0, wx, wy, lane);
}

// ------------------------------------------------------------------------------------------------------------------------
// dot_conv_kernel<PK, AU> -- the same engine on the packed dot-product instructions, for layers whose weights fit 8 bits
// (every reference config; FixedPoint weights of 9..16 bits stay on imad_conv_kernel):
//   PK = 2 (IDP.2A): activation lanes of 9..16 bits; a patch word holds 2 channels, a weight word holds
//                    {w(l,c), w(l,c+1), w(l+32,c), w(l+32,c+1)}: dp2a.lo feeds channel l, dp2a.hi channel l+32;
//   PK = 4 (IDP.4A): activation lanes of <= 8 bits; a patch word holds 4 channels, one weight word per output channel.
// The weight table is laid out on the host in the order the inner loop reads it (dot_pack_weights below), so staging a channel
// chunk is a straight 16-byte copy; stream words whose lanes are exactly 32 / PK bits wide are copied without unpacking.
// One LDS.128 of the patch now carries 4*PK channels of a pixel, so a lane issues 16 + PK/2 shared loads per 128 dot
// instructions = 128*PK MACs, and the patch takes 4/PK bytes per (pixel, channel): a 64-channel 3x3 layer is one channel chunk.
// AU: the activation lanes are unsigned (dp*.u32.s32); the sums are exact in 32 bits either way, as with IMAD.
template <int PK, bool AU, bool HI>
__device__ __forceinline__ int32_t dotacc(uint32_t a, uint32_t b, int32_t c) {
  int32_t d;
  if (PK == 4) {
    if (AU) asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  } else if (HI) {
    if (AU) asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  } else {
    if (AU) asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  }
  return d;
}

// Patch layout: uint4 [group][pixel]: the 16 patch reads of a (tap, group) step differ only by a per-row register (the warp's four
// rows) plus a compile-time column offset (SX = the horizontal stride; 0 = run-time value), and the (tap, group) part of the address
// is a block-uniform add -- no integer multiply-add on the address path: IMAD shares the pipe the dot products run on.
template <int PK, bool AU, int SX, bool DECONV, int PKY, int PKX>
__device__ __forceinline__ void dot_taps(const DirectParams& p, const uint8_t* __restrict__ patch, const uint4* __restrict__ wsm, int ngrp, int lane,
                                         const int (&rowoff)[PXB], int32_t (&acc)[PXB * PXB][2]) {
  constexpr int H = PK == 4 ? 2 : 1;  // weight quads per group: PK = 4 keeps channels l and l+32 in separate words
  const int gbytes = p.patch_h * p.patch_w * 16, gstride = p.KX * p.KY * H * 32;
  const int colstep = (SX ? SX : p.SXe) * 16;
  for (int ky = DECONV ? PKY : 0; ky < p.KY; ky += DECONV ? 2 : 1)
    for (int kx = DECONV ? PKX : 0; kx < p.KX; kx += DECONV ? 2 : 1) {
      const uint4* wrow = wsm + (size_t)(ky * p.KX + kx) * H * 32 + lane;
      const uint8_t* ptap = patch + (ky * p.DY * p.patch_w + kx * p.DX) * 16;
      for (int g = 0; g < ngrp; g++) {
        const uint4 w0 = wrow[(size_t)g * gstride];
        const uint4 w1 = PK == 4 ? wrow[(size_t)g * gstride + 32] : w0;
        const uint8_t* pg = ptap + g * gbytes;
#pragma unroll
        for (int i = 0; i < PXB * PXB; i++) {
          const int ly = i >> 2, lx = i & 3;
          if (DECONV && (((ly ^ PKY) & 1) || ((lx ^ PKX) & 1))) continue;  // compile-time: structural zero of the inserted frame
          const uint4 a = *reinterpret_cast<const uint4*>(pg + rowoff[ly] + lx * colstep);
          int32_t s0 = acc[i][0], s1 = acc[i][1];
          s0 = dotacc<PK, AU, false>(a.x, w0.x, s0); s1 = dotacc<PK, AU, true>(a.x, w1.x, s1);
          s0 = dotacc<PK, AU, false>(a.y, w0.y, s0); s1 = dotacc<PK, AU, true>(a.y, w1.y, s1);
          s0 = dotacc<PK, AU, false>(a.z, w0.z, s0); s1 = dotacc<PK, AU, true>(a.z, w1.z, s1);
          s0 = dotacc<PK, AU, false>(a.w, w0.w, s0); s1 = dotacc<PK, AU, true>(a.w, w1.w, s1);
          acc[i][0] = s0; acc[i][1] = s1;
        }
      }
    }
}

template <int PK, bool AU, int SX>
__global__ void __launch_bounds__(256, 2) dot_conv_kernel(const DirectParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int G = 4 * PK, H = PK == 4 ? 2 : 1;  // channels per group (one LDS.128 of the patch)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int img = blockIdx.z;
  const int chb = blockIdx.y * CH_PER_CTA, ch0 = chb + lane;
  const int ox0 = tx * TX, oy0 = ty * TY;
  const int wx = (warp & 3) * PXB, wy = (warp >> 2) * PXB;
  const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;
  int32_t acc[PXB * PXB][2];
#pragma unroll
  for (int i = 0; i < PXB * PXB; i++) acc[i][0] = acc[i][1] = 0;
  const int vx0 = ox0 * p.SXe, vy0 = oy0 * p.SYe;
  uint32_t* patch = reinterpret_cast<uint32_t*>(smem_raw);
  int rowoff[PXB];  // byte offset of the warp's four pixel rows inside a group plane of the patch
#pragma unroll
  for (int ly = 0; ly < PXB; ly++) rowoff[ly] = ((wy + ly) * p.SYe * p.patch_w + wx * p.SXe) * 16;
  const bool word_copy = p.in_bits * PK == 32 && p.in_word_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0;
  for (int c0 = 0; c0 < p.C; c0 += p.CC) {
    const int cc = min(p.CC, p.C - c0), ngrp = (cc + G - 1) / G, wpp = ngrp * 4;
    const int npix = p.patch_h * p.patch_w;
    uint32_t* wsm = patch + (size_t)npix * wpp;
    // ---- stage the patch, PK channels per word: zero padding / zero insertion resolved here; channels past cc are zeros
    const int total = npix * wpp;
    for (int idx = tid; idx < total; idx += blockDim.x) {  // (global reads walk a pixel's words; the store scatters them by group)
      const int w = idx % wpp, pix = idx / wpp;
      const int px = pix % p.patch_w, py = pix / p.patch_w;
      int sx, sy;
      const bool okx = map_coord(vx0 + px, p.PAD, p.deconv, p.IX, &sx);
      const bool oky = map_coord(vy0 + py, p.PADY, p.deconv, p.IY, &sy);
      uint32_t v = 0;
      if (okx && oky && w * PK < cc) {
        const uint8_t* word = in + ((size_t)sy * p.IX + sx) * p.in_word_bytes;
        if (word_copy && (w + 1) * PK <= cc) {  // lanes of exactly 32 / PK bits: the stream word already is the packed operand
          v = __ldg(reinterpret_cast<const uint32_t*>(word) + c0 / PK + w);
        } else {
#pragma unroll
          for (int j = 0; j < PK; j++) {
            const int c = w * PK + j;
            if (c < cc) v |= ((uint32_t)load_lane_any(word, c0 + c, p.in_bits, p.in_signed) & (PK == 4 ? 0xFFu : 0xFFFFu)) << (j * (32 / PK));
          }
        }
      }
      patch[(((w >> 2) * npix + pix) << 2) + (w & 3)] = v;
    }
    // ---- stage the weights of this channel chunk: the table is stored in the order the inner loop reads it,
    // uint4 [channel block][group][tap][H][lane] (fcb_api.cu), so a chunk is one contiguous run
    const int wtotal = ngrp * p.KX * p.KY * H * 32;
    const uint4* wsrc = reinterpret_cast<const uint4*>(p.wt) + ((size_t)blockIdx.y * ((p.C + G - 1) / G) + c0 / G) * (p.KX * p.KY * H * 32);
    uint4* wdst = reinterpret_cast<uint4*>(wsm);
    for (int idx = tid; idx < wtotal; idx += blockDim.x) wdst[idx] = __ldg(wsrc + idx);
    __syncthreads();
    const uint4* wsm4 = reinterpret_cast<const uint4*>(wsm);
    const uint8_t* pb = reinterpret_cast<const uint8_t*>(patch);
    if (p.deconv) {
      dot_taps<PK, AU, SX, true, 0, 0>(p, pb, wsm4, ngrp, lane, rowoff, acc);
      dot_taps<PK, AU, SX, true, 0, 1>(p, pb, wsm4, ngrp, lane, rowoff, acc);
      dot_taps<PK, AU, SX, true, 1, 0>(p, pb, wsm4, ngrp, lane, rowoff, acc);
      dot_taps<PK, AU, SX, true, 1, 1>(p, pb, wsm4, ngrp, lane, rowoff, acc);
    } else {
      dot_taps<PK, AU, SX, false, 0, 0>(p, pb, wsm4, ngrp, lane, rowoff, acc);
    }
    __syncthreads();
  }
  direct_epilogue(p, acc, img, ch0, ox0, oy0, wx, wy, lane);
}

// Host side: weights [OFM][K = taps * C] (tap-major, as everywhere in this library) -> the table dot_conv_kernel<pk> stages,
// uint32 [OFMp / 64][ceil(C / G)][taps][H][32 lanes][4]: PK = 2: bytes {w(l,c), w(l,c+1), w(l+32,c), w(l+32,c+1)};
// PK = 4: bytes w(l + 32 h, c .. c+3).
std::vector<uint32_t> dot_pack_weights(const std::vector<int32_t>& W, int OFM, int OFMp, int C, int taps, int pk) {
  const int G = 4 * pk, H = pk == 4 ? 2 : 1, ngrp = (C + G - 1) / G, K = taps * C;
  std::vector<uint32_t> T((size_t)(OFMp / 64) * ngrp * taps * H * 128, 0u);
  auto w = [&](int ch, int tap, int c) -> uint32_t {
    return ch < OFM && c < C ? (uint32_t)(uint8_t)(int8_t)W[(size_t)ch * K + (size_t)tap * C + c] : 0u;
  };
  size_t o = 0;
  for (int blk = 0; blk < OFMp / 64; blk++)
    for (int g = 0; g < ngrp; g++)
      for (int tap = 0; tap < taps; tap++)
        for (int h = 0; h < H; h++)
          for (int l = 0; l < 32; l++)
            for (int j = 0; j < 4; j++, o++) {
              uint32_t v = 0;
              for (int b = 0; b < 4; b++) {
                const int c = g * G + j * pk + (pk == 4 ? b : (b & 1));
                const int ch = blk * 64 + l + (pk == 4 ? 32 * h : 32 * (b >> 1));
                v |= w(ch, tap, c) << (8 * b);
              }
              T[o] = v;
            }
  return T;
}

// channels per chunk (0: does not fit) and shared-memory bytes of dot_conv_kernel<pk>
int dot_chunk_channels(int patch_w, int patch_h, int taps, int C, int pk, size_t budget) {
  const int G = 4 * pk;
  const size_t per_group = (size_t)patch_w * patch_h * 16 + (size_t)taps * (pk == 4 ? 2 : 1) * 512;
  const int groups = (int)(budget / per_group);
  if (groups < 1) return 0;
  return std::min(groups * G, (C + G - 1) / G * G);
}
size_t dot_smem_bytes(int patch_w, int patch_h, int taps, int cc, int pk) {
  const int G = 4 * pk;
  const size_t ngrp = (size_t)(cc + G - 1) / G;
  return ngrp * ((size_t)patch_w * patch_h * 16 + (size_t)taps * (pk == 4 ? 2 : 1) * 512);
}

size_t direct_smem_bytes(int engine, int patch_w, int patch_h, int cc) {
  (void)engine;
  return (size_t)patch_w * patch_h * cc * 4;
}
// imad_conv_kernel: patch (channel pitch rounded up to 4) + packed weight pairs of the chunk
size_t imad_smem_bytes(int patch_w, int patch_h, int taps, int cc) {
  const size_t cc4 = (size_t)((cc + 3) & ~3);
  return (size_t)patch_w * patch_h * cc4 * 4 + (size_t)taps * cc4 * 128;
}

int launch_direct(const DirectParams& p, int engine, int n_images, size_t smem_bytes, cudaStream_t st) {
  dim3 grid(p.tiles_x * p.tiles_y, (p.OFM + CH_PER_CTA - 1) / CH_PER_CTA, 1);
  // grid.z is limited to 65535 images per launch
  for (int n0 = 0; n0 < n_images; n0 += 65535) {
    DirectParams q = p;
    const int nb = n_images - n0 < 65535 ? n_images - n0 : 65535;
    q.in = p.in + (size_t)n0 * p.in_img_bytes;
    q.out = p.out + (size_t)n0 * p.out_img_bytes;
    grid.z = nb;
    if (engine == ENG_XNOR) {
      FCB_CUDA_OK(cudaFuncSetAttribute(direct_conv_kernel<ENG_XNOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      direct_conv_kernel<ENG_XNOR><<<grid, 256, smem_bytes, st>>>(q);
    } else if (p.mul_kind == FCB_W_BINARY_XNOR) {  // xnor layers the popcount engine cannot take (IFM_CH % 32 != 0): a == w per lane
      FCB_CUDA_OK(cudaFuncSetAttribute(direct_conv_kernel<ENG_IMAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      direct_conv_kernel<ENG_IMAD><<<grid, 256, smem_bytes, st>>>(q);
    } else if (p.dot_pack) {
      void (*k)(DirectParams) = nullptr;
      const int sx = p.SXe <= 2 ? p.SXe : 0, sel = (p.dot_pack == 4 ? 6 : 0) + (p.in_signed ? 0 : 3) + sx;
      switch (sel) {  // {IDP.2A, IDP.4A} x {signed, unsigned lanes} x {stride generic, 1, 2}
        case 0: k = dot_conv_kernel<2, false, 0>; break;
        case 1: k = dot_conv_kernel<2, false, 1>; break;
        case 2: k = dot_conv_kernel<2, false, 2>; break;
        case 3: k = dot_conv_kernel<2, true, 0>; break;
        case 4: k = dot_conv_kernel<2, true, 1>; break;
        case 5: k = dot_conv_kernel<2, true, 2>; break;
        case 6: k = dot_conv_kernel<4, false, 0>; break;
        case 7: k = dot_conv_kernel<4, false, 1>; break;
        case 8: k = dot_conv_kernel<4, false, 2>; break;
        case 9: k = dot_conv_kernel<4, true, 0>; break;
        case 10: k = dot_conv_kernel<4, true, 1>; break;
        default: k = dot_conv_kernel<4, true, 2>; break;
      }
      FCB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      k<<<grid, 256, smem_bytes, st>>>(q);
    } else {
      FCB_CUDA_OK(cudaFuncSetAttribute(imad_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      imad_conv_kernel<<<grid, 256, smem_bytes, st>>>(q);
    }
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
}

}  // namespace fcb
