// Baseline-JPEG decode on the GPU: row N2 of the hot-path scope table (image ingest).
//
// Reference semantics: `Image.open(path).convert("RGB")` (/root/reference/main.py:330-334 load_image; callers main.py:165, 412),
// i.e. Pillow on libjpeg-turbo with its defaults (JDCT_ISLOW, fancy upsampling, YCbCr -> RGB).  The result is BIT-IDENTICAL to
// Pillow's for every file inside the envelope below (oracle: oracle/jpeg_ref.py, pinned against Pillow itself):
//
//   envelope   SOF0 / SOF1 (sequential Huffman, 8-bit), one interleaved scan, 1 component (grayscale) or 3 components YCbCr with
//              luma sampling 1x1 (4:4:4), 2x1 (4:2:2) or 2x2 (4:2:0) and 1x1 chroma, Huffman table ids 0 / 1, restart intervals.
//              Everything else (progressive, arithmetic, CMYK, RGB-coded, 4:4:0 ...) is reported per file by the plan
//              (IIC_JPEG_UNSUPPORTED) and stays on the caller's host path, as PNG / URL inputs do.
//
// Stages (one launch each for the whole batch, all on the caller's stream):
//   host   iic_jpeg_plan_create   marker parsing (jdmarker.c), Huffman look-up tables (jdhuff.c jpeg_make_d_derived_tbl), layout
//   1      jpeg_huffman_kernel    entropy decoding is inherently serial inside a scan, so the parallelism is ACROSS the images of the
//                                 batch: one warp per image, lane 0 walks the bit stream (64-bit accumulator, 4 bytes per refill
//                                 when no 0xFF is among them; one shared-memory look-up per AC symbol for code + value bits <= 11, canonical-code walk otherwise), the
//                                 whole warp writes every finished 8x8 block (still in zig-zag order) as one 128-byte store
//   2      jpeg_idct_kernel       dequantisation + jpeg_idct_islow (jidctint.c) incl. the range-limit table; 8 threads per block
//   3      jpeg_color_kernel      fancy (triangle) upsampling of the chroma planes (jdsample.c, context rows as jdmainct.c) and
//                                 ycc_rgb_convert (jdcolor.c) -> uint8 HWC RGB at the image's own size, one thread per pixel
// The decoded pixels never exist on the host: iic_preprocess reads them where they were written.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/iic.h"
#include "kernels.h"

namespace iic {

namespace {

constexpr int kLutBits = 10;
constexpr int kLutSize = 1 << kLutBits;

struct HuffTable {            // device view of one Huffman table
  uint16_t lut[kLutSize];     // (code length << 8) | symbol for codes of <= kLutBits bits; 0: longer code
  int32_t maxcode[18];        // largest code of each length (as an integer of that many bits), -1: none
  int32_t valoff[18];         // index of the first symbol of that length minus its smallest code
  uint8_t vals[256];
};
static_assert(sizeof(HuffTable) % 8 == 0, "tables are laid out back to back");

// AC symbols whose code AND value bits fit the look-ahead (code length + size <= kFastBits) decode with ONE look-up and no
// data-dependent branch:  entry = value << 16 | advance << 8 | total bits  (0: not covered), where `advance` is what the symbol
// adds to the zig-zag position: run + 1 for a coefficient, 16 for ZRL, 64 for EOB (ends the block through the loop condition).
// One table per distinct AC table of the batch; a warp copies the two of its image into shared memory.
constexpr int kFastBits = 11;
constexpr int kFastSize = 1 << kFastBits;
struct FastAc {
  int32_t e[kFastSize];
};

struct alignas(16) JpegComp {
  uint16_t qt[64];            // quantisation table, natural order
  unsigned long long plane_off;   // byte offset of the sample plane in the scratch buffer
  int blk0, pad0_;            // index of the component's first block inside an MCU
  int h, v;                   // sampling factors (1 x 1 for a lone component)
  int bw, bh;                 // block grid (padded to whole MCUs)
  int dw, dh;                 // real samples of the component (downsampled_width / height)
  int pitch;                  // bytes per plane row = 8 bw
  int td, ta;                 // Huffman table slots (0 / 1)
  int pad_[3];
};
// Coefficient blocks are stored in DECODING order - [MCU][block of the MCU][64] - once per CHAIN of the image (see
// jpeg_huffman_kernel): a chain that starts in the middle of the entropy-coded segment cannot know its MCU index.
constexpr int kMaxChains = 8;
constexpr int kChainWindow = 32;   // MCUs a speculative chain decodes (and discards) before it publishes a boundary
struct ChainInfo {                 // written by the Huffman kernel, read by the IDCT kernel
  int start_mcu[kMaxChains];       // first MCU of the image the chain's stored blocks belong to
  int n_mcu[kMaxChains];           // stored MCUs that are valid (0: the chain never synchronised / is not used)
  int dc_off[kMaxChains][3];       // DC predictor of every component at the chain's first stored MCU
};
static_assert(sizeof(JpegComp) == 192, "layout shared by host and device");

struct alignas(16) JpegImage {
  unsigned long long data_off;    // first entropy-coded byte inside the blob
  unsigned char* out;             // uint8 [height][width][3]
  unsigned int data_len;          // bytes from data_off to the end of the file
  int width, height, ncomp, hmax, vmax, mcux, mcuy, restart_interval;
  short tab[4];                   // DC0, DC1, AC0, AC1: index into the batch's table array (identical tables are stored once)
  int nchains;                    // decoding chains of this image (1 or the batch's chain count)
  unsigned long long coef_off;    // coefficient blocks of chain 0; chain c at coef_off + c * chain_stride
  unsigned long long chain_stride;
  unsigned long long info_off;    // ChainInfo
  int bpm;                        // blocks per MCU
  int pad_;
  JpegComp comp[3];
};
static_assert(sizeof(JpegImage) == 96 + 3 * 192, "layout shared by host and device");

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------------------ entropy decoding
struct BitReader {
  const uint8_t* p;     // next unread byte
  const uint8_t* end;
  uint64_t acc;         // MSB-aligned bit accumulator
  int n;                // valid bits in acc
  int marker;           // a marker has been reached: zeros are fed from here on (jdhuff.c jpeg_fill_bit_buffer)
  int fake;             // zero bits fed so far (they sit below the real bits of the accumulator)
};

__device__ __forceinline__ uint32_t load_be32(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  const uint32_t lo = __ldg(w), hi = __ldg(w + 1);                  // the blob is readable 8 bytes past its end
  const uint32_t le = __funnelshift_r(lo, hi, uint32_t(a & 3) * 8);
  return __byte_perm(le, 0, 0x0123);
}

__device__ __forceinline__ void refill(BitReader& br) {
  while (br.n <= 32) {
    if (!br.marker && br.p + 4 <= br.end) {
      const uint32_t w = load_be32(br.p);
      if (((~w - 0x01010101u) & w & 0x80808080u) == 0u) {          // no 0xFF among the four bytes: no stuffing, no marker
        br.acc |= uint64_t(w) << (32 - br.n);
        br.n += 32;
        br.p += 4;
        continue;
      }
    }
    uint32_t b = 0;
    if (!br.marker && br.p < br.end) {
      b = *br.p++;
      if (b == 0xFF) {
        uint32_t b2 = br.p < br.end ? *br.p : 0xD9u;
        while (b2 == 0xFF && br.p < br.end) {                       // fill bytes
          ++br.p;
          b2 = br.p < br.end ? *br.p : 0xD9u;
        }
        ++br.p;
        if (b2 == 0) b = 0xFF;
        else { br.marker = int(b2); b = 0; }
      }
    }
    else br.fake += 8;
    br.acc |= uint64_t(b) << (56 - br.n);
    br.n += 8;
  }
}

// Canonical form of a reader state: whole buffered bytes go back to the stream, at most 7 bits of the last data byte stay
// (a data byte 0xFF occupies two raw bytes, FF 00).  Two readers that have consumed the same bits of the same stream end
// up with the same (p, n): the position below identifies a point of the stream.
__device__ __forceinline__ void normalize(BitReader& br, const uint8_t* lo) {
  while (br.n >= 8) {
    const uint8_t* q = br.p - 1;
    if (q > lo && q[0] == 0x00 && q[-1] == 0xFF) --q;
    br.p = q;
    br.n -= 8;
  }
  br.acc = br.n ? (br.acc & ~(~0ull >> br.n)) : 0ull;
}
__device__ __forceinline__ unsigned long long stream_pos(const BitReader& br, const uint8_t* base) {
  return (unsigned long long)(br.p - base) * 8ull - (unsigned long long)br.n;
}

// byte-align and step over the RSTn marker (jdhuff.c process_restart / jdmarker.c read_restart_marker)
__device__ __forceinline__ void restart(BitReader& br) {
  br.acc = 0;
  br.n = 0;
  if (!br.marker) {
    while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] != 0 && br.p[1] != 0xFF)) ++br.p;
    br.p += 2;
  }
  br.marker = 0;
}

__device__ __forceinline__ int decode_symbol(BitReader& br, const HuffTable* __restrict__ t) {
  const uint32_t e = __ldg(&t->lut[uint32_t(br.acc >> (64 - kLutBits))]);
  if (e != 0) {
    const int len = int(e >> 8);
    br.acc <<= len;
    br.n -= len;
    return int(e & 255u);
  }
  const int code16 = int(br.acc >> 48);
#pragma unroll 1
  for (int l = kLutBits + 1; l <= 16; ++l) {
    const int c = code16 >> (16 - l);
    if (c <= __ldg(&t->maxcode[l])) {
      br.acc <<= l;
      br.n -= l;
      return int(__ldg(&t->vals[(c + __ldg(&t->valoff[l])) & 255]));
    }
  }
  br.acc <<= 16;   // corrupt stream: libjpeg warns and returns 0
  br.n -= 16;
  return 0;
}

__device__ __forceinline__ int receive_extend(BitReader& br, int s) {   // HUFF_EXTEND
  const int v = int(br.acc >> (64 - s));
  br.acc <<= s;
  br.n -= s;
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// one 8x8 block: blk receives the quantised coefficients in ZIG-ZAG order (the IDCT kernel undoes it on load).
// After a refill the accumulator holds > 32 bits, enough for THREE one-look-up symbols (<= 11 bits each) before the next check;
// the common path falls through every branch (a lone warp pays ~20 cycles for each taken one).
__device__ __forceinline__ void decode_block(BitReader& br, const HuffTable* __restrict__ dc, const HuffTable* __restrict__ ac,
                                             const int32_t* __restrict__ fast, int& pred, int16_t* blk) {
  if (br.n <= 32) refill(br);
  int s = decode_symbol(br, dc) & 15;
  if (s) pred += receive_extend(br, s);
  blk[0] = int16_t(pred);
  int k = 1;
  auto fast_step = [&](int e) {
    const int len = e & 255, pos = k + ((e >> 8) & 255) - 1;
    br.acc <<= len;
    br.n -= len;
    blk[pos] = int16_t(e >> 16);     // ZRL writes a zero into a still-zero slot; EOB (and a corrupt run) lands in the 64 spare
    k = pos + 1;                     // slots behind the block, which nobody reads: no store predicate, one exit test per symbol
  };
#pragma unroll 1
  while (k < 64) {
    if (__builtin_expect(br.n <= 32, 0)) refill(br);
    int e = fast[uint32_t(br.acc >> (64 - kFastBits))];
    if (__builtin_expect((e & 255) == 0, 0)) {          // long code or large value: canonical walk + receive / extend
      const int rs = decode_symbol(br, ac);
      s = rs & 15;
      int v = 0;
      if (s) v = receive_extend(br, s);
      const int pos = k + (s ? (rs >> 4) + 1 : ((rs >> 4) == 15 ? 16 : 64)) - 1;
      blk[pos] = int16_t(v);
      k = pos + 1;
      continue;
    }
    fast_step(e);
    if (k >= 64) break;
    e = fast[uint32_t(br.acc >> (64 - kFastBits))];
    if (__builtin_expect((e & 255) == 0, 0)) continue;
    fast_step(e);
    if (k >= 64) break;
    e = fast[uint32_t(br.acc >> (64 - kFastBits))];
    if (__builtin_expect((e & 255) == 0, 0)) continue;
    fast_step(e);
  }
}

// Entropy decoding is serial inside a scan, so the parallelism comes from two places:
//   ACROSS the images of the batch: a CTA of kHuffWarps warps serves kHuffWarps / chains images, one warp per CHAIN.  The
//       one-look-up tables of the CTA's first image sit in shared memory; the host orders the images by AC table pair, so the
//       other warps nearly always read them there too (else from global memory / L1).
//   INSIDE an image (chains > 1: batches too small to fill the GPU with one chain per image): chain c starts at byte
//       c * len / chains of the entropy-coded segment in a guessed state.  Huffman streams self-synchronise: after `window` MCUs
//       (decoded and discarded) chain c brings its reader into canonical form, PUBLISHES that stream position, zeroes its DC
//       predictors and from there on stores blocks under its own MCU count.  Chain c - 1 (whose state is true by induction
//       from chain 0) compares its canonical position at every MCU boundary behind the successor's start byte: equal -> both
//       are in the same state from there on, so it stops and the successor's blocks are the image's; past it without a match
//       -> the successor never synchronised in its window, is told to stop, and the chain carries on towards the next one.
//       Every outcome yields Pillow's pixels; what varies is who decoded them.
// lane 0 of a warp walks the bit stream; the whole warp writes each finished block (zig-zag order) as one 128-byte store.
constexpr int kHuffWarps = 8;
static_assert(kHuffWarps == kMaxChains, "one warp per chain");

struct ChainShared {
  unsigned long long pub[kMaxChains];   // published position of the chain's first stored MCU; ~0: not yet; 0: never
  int dead[kMaxChains];
  int n_mcu[kMaxChains], matched[kMaxChains], end_pred[kMaxChains][3];
};

__global__ void __launch_bounds__(32 * kHuffWarps) jpeg_huffman_kernel(const JpegImage* __restrict__ imgs, const int* __restrict__ order,
                                                                       int n_images, const HuffTable* __restrict__ tables,
                                                                       const FastAc* __restrict__ fast_tables,
                                                                       const uint8_t* __restrict__ blob, uint8_t* __restrict__ scratch,
                                                                       int chains, int window) {
  __shared__ __align__(16) int16_t blk_s[kHuffWarps][128];   // a block + 64 spare slots (decode_block: positions up to 126)
  __shared__ __align__(16) int32_t fast_s[2][kFastSize];
  __shared__ ChainShared sh_all[kHuffWarps];               // one per image of the CTA
  struct BlockDesc {                                       // what block `bi` of an MCU decodes with (per warp: images differ)
    const HuffTable* dc;
    const HuffTable* ac;
    const int32_t* fast;                                   // the image's own table in global memory ...
    int comp, ta;                                          // ... or fast_s[ta] when it equals the CTA's shared copy
  };
  __shared__ BlockDesc bd_s[kHuffWarps][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ipc = kHuffWarps / chains;                     // images per CTA
  const int first = int(blockIdx.x) * ipc;
  const JpegImage& im0 = imgs[order[first]];
  for (int t = 0; t < 2; ++t) {
    const uint4* src = reinterpret_cast<const uint4*>(fast_tables[im0.tab[2 + t]].e);
    uint4* dst = reinterpret_cast<uint4*>(fast_s[t]);
    for (int i = threadIdx.x; i < kFastSize / 4; i += 32 * kHuffWarps) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x < kHuffWarps * kMaxChains) {
    ChainShared& z = sh_all[threadIdx.x / kMaxChains];
    const int c = threadIdx.x % kMaxChains;
    z.pub[c] = ~0ull;
    z.dead[c] = 0;
    z.n_mcu[c] = 0;
    z.matched[c] = -1;
    z.end_pred[c][0] = z.end_pred[c][1] = z.end_pred[c][2] = 0;
  }
  __syncthreads();
  const int img_slot = first + warp / chains;
  const int chain = warp % chains;
  ChainShared& sh = sh_all[warp / chains];
  const bool active = img_slot < n_images && chain < imgs[order[img_slot < n_images ? img_slot : first]].nchains;
  if (active) {
    const JpegImage& im = imgs[order[img_slot]];
    const int nchains = im.nchains;
    int16_t* blk = blk_s[warp];
    const int32_t* fast[2];
    for (int t = 0; t < 2; ++t) fast[t] = im.tab[2 + t] == im0.tab[2 + t] ? fast_s[t] : fast_tables[im.tab[2 + t]].e;
    reinterpret_cast<uint32_t*>(blk)[lane] = 0u;
    __syncwarp();
    const uint8_t* seg = blob + im.data_off;
    BitReader br;
    br.p = seg + (size_t(im.data_len) * size_t(chain)) / size_t(nchains);
    br.end = seg + im.data_len;
    if (chain > 0 && br.p[-1] == 0xFF && br.p[0] == 0x00) ++br.p;      // not on the stuffing byte of an FF 00 pair
    br.acc = 0;
    br.n = 0;
    br.marker = 0;
    br.fake = 0;
    BlockDesc* bd = bd_s[warp];
    if (lane == 0) {
      int bi = 0;
      for (int ci = 0; ci < im.ncomp; ++ci)
        for (int b = 0; b < im.comp[ci].h * im.comp[ci].v && bi < 12; ++b, ++bi) {
          bd[bi].dc = tables + im.tab[im.comp[ci].td];
          bd[bi].ac = tables + im.tab[2 + im.comp[ci].ta];
          bd[bi].fast = fast[im.comp[ci].ta];
          bd[bi].comp = ci;
          bd[bi].ta = im.comp[ci].ta;
        }
    }
    __syncwarp();
    int pred0 = 0, pred1 = 0, pred2 = 0;
    int todo = im.restart_interval;
    const int total = im.mcux * im.mcuy, bpm = im.bpm;
    int16_t* region = reinterpret_cast<int16_t*>(scratch + im.coef_off + im.chain_stride * size_t(chain));
    int spec = chain > 0 ? window : 0;             // MCUs still to decode before this chain's blocks count (every lane counts)
    int stored = 0, decoded = 0;
    uint32_t* out_blk = reinterpret_cast<uint32_t*>(region) + lane;      // this lane's word of the next stored block
    int cand = chain + 1;                          // the successor this chain expects to meet
    const uint8_t* cand_start = seg + (size_t(im.data_len) * size_t(cand)) / size_t(nchains);
    int matched = -1;
#pragma unroll 1
    for (;;) {
      int stop = 0;
      if (lane == 0) {
        if (decoded >= total) stop = 1;
        if (nchains > 1 && !stop) {
          if (chain > 0 && spec == 0 && stored == 0) {                  // end of the window: first trusted boundary
            if (br.marker) {
              stop = 1;                                                  // ran into the end of the data inside the window
            } else {
              normalize(br, seg);
              pred0 = pred1 = pred2 = 0;
              sh.pub[chain] = stream_pos(br, seg);
              __threadfence_block();
            }
          }
          if (!stop && *reinterpret_cast<volatile int*>(&sh.dead[chain])) stop = 1;
          while (!stop && cand < nchains && br.p >= cand_start && !br.marker) {
            normalize(br, seg);
            const unsigned long long pos = stream_pos(br, seg);
            unsigned long long pb;
            while ((pb = *reinterpret_cast<volatile unsigned long long*>(&sh.pub[cand])) == ~0ull) __nanosleep(200);
            if (spec == 0 && pos == pb) {
              matched = cand;
              stop = 1;
            } else if (pos > pb) {                                       // behind its first boundary without meeting it
              *reinterpret_cast<volatile int*>(&sh.dead[cand]) = 1;
              ++cand;
              cand_start = seg + (size_t(im.data_len) * size_t(cand)) / size_t(nchains);
            } else {
              break;                                                     // not there yet
            }
          }
          if (!stop && chain > 0 && br.fake > 0 && br.n - br.fake < 8) stop = 1;   // all real bits consumed
        }
        if (!stop && im.restart_interval) {
          if (todo == 0) {
            restart(br);
            pred0 = pred1 = pred2 = 0;
            todo = im.restart_interval;
          }
          --todo;
        }
      }
      stop = __shfl_sync(0xffffffffu, stop, 0);
      if (stop) break;
      const bool keep = spec == 0;
#pragma unroll 1
      for (int bi = 0; bi < bpm; ++bi) {
        if (lane == 0) {
          const BlockDesc d = bd[bi];
          int pr = d.comp == 0 ? pred0 : (d.comp == 1 ? pred1 : pred2);
          // the shared-memory copy is addressed as such (LDS with a 32-bit address instead of a generic load)
          if (d.fast == fast_s[d.ta]) decode_block(br, d.dc, d.ac, fast_s[d.ta], pr, blk);
          else decode_block(br, d.dc, d.ac, d.fast, pr, blk);
          if (d.comp == 0) pred0 = pr; else if (d.comp == 1) pred1 = pr; else pred2 = pr;
        }
        __syncwarp();
        if (keep) {
          *out_blk = reinterpret_cast<uint32_t*>(blk)[lane];
          out_blk += 32;
        }
        reinterpret_cast<uint32_t*>(blk)[lane] = 0u;
        __syncwarp();
      }
      ++decoded;
      if (spec > 0) --spec; else ++stored;
    }
    if (lane == 0) {
      if (chain > 0 && *reinterpret_cast<volatile unsigned long long*>(&sh.pub[chain]) == ~0ull) sh.pub[chain] = 0ull;   // never got there
      sh.n_mcu[chain] = stored;
      sh.matched[chain] = matched;
      sh.end_pred[chain][0] = pred0;
      sh.end_pred[chain][1] = pred1;
      sh.end_pred[chain][2] = pred2;
      __threadfence_block();
    }
  }
  __syncthreads();
  if (int(threadIdx.x) < ipc && first + int(threadIdx.x) < n_images) {   // who decoded which MCUs: follow the matches from chain 0
    const JpegImage& im = imgs[order[first + threadIdx.x]];
    const ChainShared& z = sh_all[threadIdx.x];
    ChainInfo* info = reinterpret_cast<ChainInfo*>(scratch + im.info_off);
    const int total = im.mcux * im.mcuy;
    for (int c = 0; c < kMaxChains; ++c) {
      info->start_mcu[c] = 0;
      info->n_mcu[c] = 0;
      info->dc_off[c][0] = info->dc_off[c][1] = info->dc_off[c][2] = 0;
    }
    int start = 0, dc[3] = {0, 0, 0};
    for (int c = 0; c >= 0 && c < kMaxChains && start < total;) {
      const int n = min(z.n_mcu[c], total - start);
      info->start_mcu[c] = start;
      info->n_mcu[c] = n;
      for (int k = 0; k < 3; ++k) info->dc_off[c][k] = dc[k];
      for (int k = 0; k < 3; ++k) dc[k] += z.end_pred[c][k];
      start += n;
      c = z.matched[c];
    }
  }
}

// ------------------------------------------------------------------------------------------------ dequantisation + IDCT
// one pass of jpeg_idct_islow over 8 values (jidctint.c; 13-bit constants)
__device__ __forceinline__ void idct8(const int (&in)[8], int (&out)[8], int descale) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137);
  int tmp3 = z1 + z2 * 6270;
  z2 = in[0];
  z3 = in[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7];
  tmp1 = in[5];
  tmp2 = in[3];
  tmp3 = in[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446;
  tmp1 *= 16819;
  tmp2 *= 25172;
  tmp3 *= 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 = z3 * (-16069) + z5;
  z4 = z4 * (-3196) + z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  const int rnd = 1 << (descale - 1);
  out[0] = (tmp10 + tmp3 + rnd) >> descale;
  out[7] = (tmp10 - tmp3 + rnd) >> descale;
  out[1] = (tmp11 + tmp2 + rnd) >> descale;
  out[6] = (tmp11 - tmp2 + rnd) >> descale;
  out[2] = (tmp12 + tmp1 + rnd) >> descale;
  out[5] = (tmp12 - tmp1 + rnd) >> descale;
  out[3] = (tmp13 + tmp0 + rnd) >> descale;
  out[4] = (tmp13 - tmp0 + rnd) >> descale;
}

// post-IDCT range limit: sample_range_limit[(x & RANGE_MASK) + CENTERJSAMPLE] of jdmaster.c prepare_range_limit_table
__device__ __forceinline__ uint32_t range_limit(int x) {
  const int i = x & 1023;
  return uint32_t(i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896)));
}

__device__ __forceinline__ int find_unit(const int* __restrict__ start, int n, int cta) {   // last u with start[u] <= cta
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&start[mid]) <= cta) lo = mid; else hi = mid;
  }
  return lo;
}

// units = (image, component) pairs; cta_start[u] = first CTA of unit u (32 blocks per CTA), cta_start[n_units] = grid size
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegImage* __restrict__ imgs, const int2* __restrict__ units,
                                                        const int* __restrict__ cta_start, int n_units,
                                                        uint8_t* __restrict__ scratch) {
  __shared__ int ws[32][8][9];
  __shared__ __align__(16) int16_t raw[32][64];            // the CTA's 32 blocks as stored: zig-zag order
  __shared__ uint8_t unzig[64];                           // natural index -> zig-zag position
  if (threadIdx.x < 64) unzig[c_zigzag[threadIdx.x]] = uint8_t(threadIdx.x);
  const int u = find_unit(cta_start, n_units, int(blockIdx.x));
  const int2 un = units[u];
  const JpegImage& im = imgs[un.x];
  const JpegComp& cp = im.comp[un.y];
  const int lb = threadIdx.x >> 3, r = threadIdx.x & 7;
  const long long nblocks = (long long)cp.bw * cp.bh;
  const long long b = (long long)(int(blockIdx.x) - __ldg(&cta_start[u])) * 32 + lb;
  const bool valid = b < nblocks;
  {
    int4 in = make_int4(0, 0, 0, 0);
    int dc_off = 0;
    if (valid) {
      // the block's place in decoding order: MCU m, block `bi` of the MCU; then the chain that decoded MCU m
      const int gy = int(b / cp.bw), gx = int(b - (long long)gy * cp.bw);
      const int m = (gy / cp.v) * im.mcux + gx / cp.h;
      const int bi = cp.blk0 + (gy % cp.v) * cp.h + gx % cp.h;
      const ChainInfo* info = reinterpret_cast<const ChainInfo*>(scratch + im.info_off);
      for (int c = 0; c < im.nchains; ++c) {
        const int s0 = __ldg(&info->start_mcu[c]), n = __ldg(&info->n_mcu[c]);
        if (m >= s0 && m < s0 + n) {
          in = *reinterpret_cast<const int4*>(scratch + im.coef_off + im.chain_stride * size_t(c) +
                                              (size_t(m - s0) * im.bpm + bi) * 128 + r * 16);
          dc_off = info->dc_off[c][un.y];
          break;
        }
      }
    }
    if (r == 0) {                                         // DC: stored relative to the chain's first MCU
      int16_t* h16 = reinterpret_cast<int16_t*>(&in);
      h16[0] = int16_t(int(h16[0]) + dc_off);
    }
    *reinterpret_cast<int4*>(&raw[lb][r * 8]) = in;
    __syncthreads();
    const uint4 q = *reinterpret_cast<const uint4*>(cp.qt + r * 8);
    const uint16_t* q16 = reinterpret_cast<const uint16_t*>(&q);
#pragma unroll
    for (int j = 0; j < 8; ++j) ws[lb][r][j] = int(raw[lb][unzig[r * 8 + j]]) * int(q16[j]);
  }
  __syncthreads();
  {
    int in[8], out[8];                                  // pass 1: column r of the block
#pragma unroll
    for (int i = 0; i < 8; ++i) in[i] = ws[lb][i][r];
    idct8(in, out, 13 - 2);
#pragma unroll
    for (int i = 0; i < 8; ++i) ws[lb][i][r] = out[i];
  }
  __syncthreads();
  {
    int in[8], out[8];                                  // pass 2: row r
#pragma unroll
    for (int j = 0; j < 8; ++j) in[j] = ws[lb][r][j];
    idct8(in, out, 13 + 2 + 3);
    if (valid) {
      const uint32_t lo = range_limit(out[0]) | (range_limit(out[1]) << 8) | (range_limit(out[2]) << 16) | (range_limit(out[3]) << 24);
      const uint32_t hi = range_limit(out[4]) | (range_limit(out[5]) << 8) | (range_limit(out[6]) << 16) | (range_limit(out[7]) << 24);
      const int by = int(b / cp.bw), bx = int(b - (long long)by * cp.bw);
      *reinterpret_cast<uint2*>(scratch + cp.plane_off + size_t(by * 8 + r) * cp.pitch + bx * 8) = make_uint2(lo, hi);
    }
  }
}

// ------------------------------------------------------------------------------------------------ upsampling + colour
__device__ __forceinline__ int upsample(const uint8_t* __restrict__ plane, const JpegComp& cp, int hmax, int vmax, int x, int y) {
  if (cp.h == hmax && cp.v == vmax) return plane[size_t(y) * cp.pitch + x];
  const int cx = x >> 1;
  if (cp.v == vmax) {                                   // h2v1
    const uint8_t* row = plane + size_t(y) * cp.pitch;
    const int p = row[cx];
    if (cp.dw <= 2) return p;                           // jinit_upsampler: box replication for very narrow components
    const int nb = (x & 1) ? min(cx + 1, cp.dw - 1) : max(cx - 1, 0);
    return (3 * p + row[nb] + ((x & 1) ? 2 : 1)) >> 2;
  }
  const int cy = y >> 1;                                // h2v2
  const uint8_t* r0 = plane + size_t(cy) * cp.pitch;
  if (cp.dw <= 2) return r0[cx];
  const int oy = (y & 1) ? min(cy + 1, cp.dh - 1) : max(cy - 1, 0);   // context rows replicate the first / last real row
  const uint8_t* r1 = plane + size_t(oy) * cp.pitch;
  const int nb = (x & 1) ? min(cx + 1, cp.dw - 1) : max(cx - 1, 0);
  const int cs = 3 * r0[cx] + r1[cx], cn = 3 * r0[nb] + r1[nb];
  return (3 * cs + cn + ((x & 1) ? 7 : 8)) >> 4;
}

__device__ __forceinline__ uint8_t clamp255(int v) { return uint8_t(min(max(v, 0), 255)); }

// the same for 4 horizontally adjacent pixels x0 .. x0 + 3 (x0 % 4 == 0): the two chroma samples they share and their two
// neighbours are read once.  Pixels beyond the image width come out as garbage and are not stored by the caller.
__device__ __forceinline__ void upsample4(const uint8_t* __restrict__ plane, const JpegComp& cp, int hmax, int vmax, int x0, int y, int w,
                                          int (&out)[4]) {
  if (cp.h == hmax && cp.v == vmax) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(plane + size_t(y) * cp.pitch + x0);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = int((v >> (8 * i)) & 255u);
    return;
  }
  if (cp.dw <= 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = upsample(plane, cp, hmax, vmax, min(x0 + i, w - 1), y);
    return;
  }
  const int c0 = x0 >> 1, cm = max(c0 - 1, 0), c1 = min(c0 + 1, cp.dw - 1), c2 = min(c0 + 2, cp.dw - 1);
  if (cp.v == vmax) {                                   // h2v1
    const uint8_t* row = plane + size_t(y) * cp.pitch;
    const int pm = row[cm], p0 = row[c0], p1 = row[c1], p2 = row[c2];
    out[0] = (3 * p0 + pm + 1) >> 2;
    out[1] = (3 * p0 + p1 + 2) >> 2;
    out[2] = (3 * p1 + p0 + 1) >> 2;
    out[3] = (3 * p1 + p2 + 2) >> 2;
    return;
  }
  const int cy = y >> 1, oy = (y & 1) ? min(cy + 1, cp.dh - 1) : max(cy - 1, 0);
  const uint8_t* r0 = plane + size_t(cy) * cp.pitch;
  const uint8_t* r1 = plane + size_t(oy) * cp.pitch;
  const int sm = 3 * r0[cm] + r1[cm], s0 = 3 * r0[c0] + r1[c0], s1 = 3 * r0[c1] + r1[c1], s2 = 3 * r0[c2] + r1[c2];
  out[0] = (3 * s0 + sm + 8) >> 4;
  out[1] = (3 * s0 + s1 + 7) >> 4;
  out[2] = (3 * s1 + s0 + 8) >> 4;
  out[3] = (3 * s1 + s2 + 7) >> 4;
}

// cta_start[i] = first CTA of image i; a thread converts 4 horizontally adjacent pixels (a CTA: 1024 pixel slots, rows padded to 4)
__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegImage* __restrict__ imgs, const int* __restrict__ image_of,
                                                         const int* __restrict__ cta_start, int n,
                                                         const uint8_t* __restrict__ scratch) {
  const int u = find_unit(cta_start, n, int(blockIdx.x));
  const JpegImage& im = imgs[image_of[u]];
  const int qpr = (im.width + 3) >> 2;                    // pixel quads per row
  const long long q = (long long)(int(blockIdx.x) - __ldg(&cta_start[u])) * 256 + threadIdx.x;
  if (q >= (long long)qpr * im.height) return;
  const int y = int(q / qpr), x0 = int(q - (long long)y * qpr) * 4;
  const int nx = min(4, im.width - x0);
  const uint8_t* yrow = scratch + im.comp[0].plane_off + size_t(y) * im.comp[0].pitch + x0;
  const uint32_t y4 = *reinterpret_cast<const uint32_t*>(yrow);      // plane rows are padded to whole blocks: always readable
  uint8_t rgb[12];
  if (im.ncomp == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = uint8_t(y4 >> (8 * i));
  } else {
    int cb4[4], cr4[4];
    upsample4(scratch + im.comp[1].plane_off, im.comp[1], im.hmax, im.vmax, x0, y, im.width, cb4);
    upsample4(scratch + im.comp[2].plane_off, im.comp[2], im.hmax, im.vmax, x0, y, im.width, cr4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int Y = int((y4 >> (8 * i)) & 255u);
      const int cb = cb4[i] - 128, cr = cr4[i] - 128;
      // jdcolor.c build_ycc_rgb_table: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554
      rgb[3 * i] = clamp255(Y + ((91881 * cr + 32768) >> 16));
      rgb[3 * i + 1] = clamp255(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
      rgb[3 * i + 2] = clamp255(Y + ((116130 * cb + 32768) >> 16));
    }
  }
  uint8_t* o = im.out + (size_t(y) * im.width + x0) * 3;
  if (nx == 4 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
#pragma unroll
    for (int w = 0; w < 3; ++w)
      o32[w] = uint32_t(rgb[4 * w]) | (uint32_t(rgb[4 * w + 1]) << 8) | (uint32_t(rgb[4 * w + 2]) << 16) | (uint32_t(rgb[4 * w + 3]) << 24);
  } else {
    for (int i = 0; i < 3 * nx; ++i) o[i] = rgb[i];
  }
}

// ------------------------------------------------------------------------------------------------ host: headers and layout
const uint8_t kZigzagHost[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HostTable {
  bool present = false;
  uint8_t bits[16];
  uint8_t vals[256];
  int count = 0;
};

struct Parsed {
  int status = IIC_JPEG_CORRUPT;
  std::string why;
  JpegImage img;               // offsets filled in by the layout pass
  HostTable tab[2][2];         // [class][id]
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void build_table(const HostTable& t, HuffTable* out) {
  std::memset(out, 0, sizeof(HuffTable));
  for (int l = 0; l < 18; ++l) out->maxcode[l] = -1;
  std::memcpy(out->vals, t.vals, size_t(t.count));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    const int cnt = t.bits[l - 1];
    out->valoff[l] = k - code;
    for (int i = 0; i < cnt; ++i, ++k, ++code) {
      if (l <= kLutBits && code < (1 << l)) {
        const int first = code << (kLutBits - l);
        for (int f = 0; f < (1 << (kLutBits - l)); ++f) out->lut[first + f] = uint16_t((l << 8) | t.vals[k]);
      }
    }
    out->maxcode[l] = cnt ? code - 1 : -1;
    code <<= 1;
  }
  out->maxcode[17] = 0x7fffffff;
}

// the one-look-up AC table (see FastAc) from the canonical look-ahead table
void build_fast_ac(const HostTable& t, FastAc* out) {
  std::memset(out, 0, sizeof(FastAc));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    for (int i = 0; i < t.bits[l - 1]; ++i, ++k, ++code) {
      const int rs = t.vals[k], r = rs >> 4, sz = rs & 15;
      const int tot = l + sz;
      if (tot > kFastBits || code >= (1 << l)) continue;
      const int adv = sz ? r + 1 : (r == 15 ? 16 : 64);
      for (int vb = 0; vb < (1 << sz); ++vb) {
        int v = 0;
        if (sz) v = vb < (1 << (sz - 1)) ? vb - (1 << sz) + 1 : vb;
        const int first = ((code << sz) | vb) << (kFastBits - tot);
        const int32_t e = int32_t((uint32_t(v) << 16) | uint32_t(adv << 8) | uint32_t(tot));
        for (int f = 0; f < (1 << (kFastBits - tot)); ++f) out->e[first + f] = e;
      }
    }
    code <<= 1;
  }
}

// marker segments up to the first SOS (jdmarker.c read_markers); fills p.img geometry, quantisation tables and Huffman tables
void parse_one(const uint8_t* d, size_t n, Parsed& p) {
  auto corrupt = [&](const char* w) { p.status = IIC_JPEG_CORRUPT; p.why = w; };
  auto unsupported = [&](const char* w) { p.status = IIC_JPEG_UNSUPPORTED; p.why = w; };
  std::memset(&p.img, 0, sizeof(p.img));
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return corrupt("not a JPEG (no SOI)");
  uint16_t qt[4][64];
  bool have_qt[4] = {false, false, false, false};
  int comp_id[3] = {0, 0, 0}, comp_tq[3] = {0, 0, 0}, comp_h[3] = {1, 1, 1}, comp_v[3] = {1, 1, 1};
  bool have_sof = false, jfif = false;
  int adobe = -1;
  size_t pos = 2;
  for (;;) {
    while (pos < n && d[pos] != 0xFF) ++pos;
    while (pos < n && d[pos] == 0xFF) ++pos;
    if (pos >= n) return corrupt("no SOS marker");
    const int m = d[pos++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) return corrupt("EOI before SOS");
    if (pos + 2 > n) return corrupt("truncated marker segment");
    const size_t seglen = (size_t(d[pos]) << 8) | d[pos + 1];
    if (seglen < 2 || pos + seglen > n) return corrupt("truncated marker segment");
    const uint8_t* seg = d + pos + 2;
    const size_t sl = seglen - 2;
    if (m == 0xDB) {
      size_t q = 0;
      while (q < sl) {
        const int pq = seg[q] >> 4, tq = seg[q] & 15;
        ++q;
        if (tq > 3 || pq > 1 || q + size_t(64 * (pq + 1)) > sl) return corrupt("bad DQT");
        for (int i = 0; i < 64; ++i) {
          qt[tq][kZigzagHost[i]] = pq ? uint16_t((seg[q] << 8) | seg[q + 1]) : uint16_t(seg[q]);
          q += size_t(pq + 1);
        }
        have_qt[tq] = true;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (have_sof) return corrupt("two SOF markers");
      if (sl < 6) return corrupt("bad SOF");
      if (seg[0] != 8) return unsupported("sample precision other than 8 bits");
      p.img.height = (seg[1] << 8) | seg[2];
      p.img.width = (seg[3] << 8) | seg[4];
      p.img.ncomp = seg[5];
      if (p.img.width <= 0 || p.img.height <= 0) return unsupported("zero image dimension (DNL)");
      // Pillow refuses files beyond 2 x MAX_IMAGE_PIXELS (179 MP) as decompression bombs; the host path keeps that decision
      if ((long long)p.img.width * p.img.height > 178956970LL) return unsupported("image larger than 178 MP");
      if (p.img.ncomp != 1 && p.img.ncomp != 3) return unsupported("component count other than 1 or 3 (CMYK / YCCK)");
      if (sl < size_t(6 + 3 * p.img.ncomp)) return corrupt("bad SOF");
      for (int c = 0; c < p.img.ncomp; ++c) {
        comp_id[c] = seg[6 + 3 * c];
        comp_h[c] = seg[7 + 3 * c] >> 4;
        comp_v[c] = seg[7 + 3 * c] & 15;
        comp_tq[c] = seg[8 + 3 * c];
        if (comp_tq[c] > 3 || comp_h[c] < 1 || comp_h[c] > 4 || comp_v[c] < 1 || comp_v[c] > 4) return corrupt("bad SOF component");
      }
      have_sof = true;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return unsupported(m == 0xC2 ? "progressive JPEG" : "lossless / hierarchical / arithmetic JPEG");
    } else if (m == 0xCC) {
      return unsupported("arithmetic coding");
    } else if (m == 0xC4) {
      size_t q = 0;
      while (q < sl) {
        if (q + 17 > sl) return corrupt("bad DHT");
        const int tc = seg[q] >> 4, th = seg[q] & 15;
        if (tc > 1 || th > 3) return corrupt("bad DHT");
        int cnt = 0;
        for (int i = 0; i < 16; ++i) cnt += seg[q + 1 + i];
        if (cnt > 256 || q + 17 + size_t(cnt) > sl) return corrupt("bad DHT");
        if (th <= 1) {
          HostTable& t = p.tab[tc][th];
          std::memcpy(t.bits, seg + q + 1, 16);
          std::memcpy(t.vals, seg + q + 17, size_t(cnt));
          t.count = cnt;
          t.present = true;
        }
        q += 17 + size_t(cnt);
      }
    } else if (m == 0xDD) {
      if (sl < 2) return corrupt("bad DRI");
      p.img.restart_interval = (seg[0] << 8) | seg[1];
    } else if (m == 0xE0) {
      if (sl >= 5 && std::memcmp(seg, "JFIF\0", 5) == 0) jfif = true;
    } else if (m == 0xEE) {
      if (sl >= 12 && std::memcmp(seg, "Adobe", 5) == 0) adobe = seg[11];
    } else if (m == 0xDA) {
      if (!have_sof) return corrupt("SOS before SOF");
      if (sl < 1 || sl < size_t(1 + 2 * seg[0] + 3)) return corrupt("bad SOS");
      const int ns = seg[0];
      if (ns != p.img.ncomp) return unsupported("multi-scan sequential JPEG");
      int hmax = 1, vmax = 1;
      for (int c = 0; c < p.img.ncomp; ++c) { hmax = std::max(hmax, comp_h[c]); vmax = std::max(vmax, comp_v[c]); }
      if (p.img.ncomp == 3) {
        // jdapimin.c default_decompress_parms: JFIF -> YCbCr; else Adobe transform 0 -> RGB; else ids 'R','G','B' -> RGB
        const bool rgb_ids = comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B';
        if (!jfif && (adobe == 0 || (adobe < 0 && rgb_ids))) return unsupported("RGB-coded JPEG");
        if (comp_h[1] != 1 || comp_v[1] != 1 || comp_h[2] != 1 || comp_v[2] != 1 ||
            !((hmax == 1 && vmax == 1) || (hmax == 2 && vmax == 1) || (hmax == 2 && vmax == 2)))
          return unsupported("chroma sampling other than 4:4:4 / 4:2:2 / 4:2:0");
      } else {
        hmax = vmax = 1;
        comp_h[0] = comp_v[0] = 1;      // a lone component is decoded block by block whatever its factors say (jdinput.c)
      }
      p.img.hmax = hmax;
      p.img.vmax = vmax;
      p.img.mcux = (p.img.width + 8 * hmax - 1) / (8 * hmax);
      p.img.mcuy = (p.img.height + 8 * vmax - 1) / (8 * vmax);
      for (int s = 0; s < ns; ++s) {
        const int cs = seg[1 + 2 * s], tt = seg[2 + 2 * s];
        if (cs != comp_id[s]) return unsupported("scan component order differs from the frame's");
        const int td = tt >> 4, ta = tt & 15;
        if (td > 1 || ta > 1) return unsupported("Huffman table id above 1");
        if (!p.tab[0][td].present || !p.tab[1][ta].present) return corrupt("scan refers to a missing Huffman table");
        if (!have_qt[comp_tq[s]]) return corrupt("component refers to a missing quantisation table");
        JpegComp& cp = p.img.comp[s];
        cp.h = comp_h[s];
        cp.v = comp_v[s];
        cp.td = td;
        cp.ta = ta;
        cp.bw = p.img.mcux * cp.h;
        cp.bh = p.img.mcuy * cp.v;
        cp.dw = (p.img.width * cp.h + hmax - 1) / hmax;
        cp.dh = (p.img.height * cp.v + vmax - 1) / vmax;
        cp.pitch = cp.bw * 8;
        std::memcpy(cp.qt, qt[comp_tq[s]], sizeof(cp.qt));
      }
      const size_t start = pos + seglen;
      p.img.data_off = start;          // relative to the file; the layout pass adds the file's offset in the blob
      p.img.data_len = unsigned(std::min<size_t>(n - start, 0xffffffffu));
      p.status = IIC_JPEG_OK;
      return;
    }
    pos += seglen;
  }
}

}  // namespace
}  // namespace iic

struct iic_jpeg_plan {
  std::vector<iic::Parsed> files;
  std::vector<iic::HuffTable> tables;  // the distinct Huffman tables of the batch, built once
  std::vector<iic::FastAc> fast;       // one-look-up companion of every table (only those of AC tables are read)
  size_t off_fast = 0;
  int chains = 1;                      // decoding chains per image (jpeg_huffman_kernel): 1 = one warp per image
  int window = iic::kChainWindow;      // MCUs a speculative chain discards before it publishes (IIC_JPEG_CHAIN_WINDOW: tests)
  std::vector<int> ok;                 // indices of the files inside the envelope, longest entropy segment first
  size_t desc_bytes = 0;               // descriptor region = staging size: images, tables, unit tables
  size_t off_tables = 0, off_order = 0, off_units = 0, off_idct_start = 0, off_color_img = 0, off_color_start = 0;
  size_t scratch_bytes = 0;
  int idct_ctas = 0, color_ctas = 0, n_units = 0;
  std::string err;
};

extern "C" {

int iic_jpeg_plan_create(const uint8_t* blob, const int64_t* offsets, int n, iic_jpeg_plan** out) {
  using namespace iic;
  if (out == nullptr) return IIC_ERR_ARG;
  *out = nullptr;
  if (blob == nullptr || offsets == nullptr || n < 0) return IIC_ERR_ARG;
  iic_jpeg_plan* pl = new (std::nothrow) iic_jpeg_plan();
  if (pl == nullptr) return IIC_ERR_ARG;
  pl->files.resize(size_t(n));
  for (int i = 0; i < n; ++i) {
    Parsed& p = pl->files[size_t(i)];
    if (offsets[i + 1] < offsets[i]) { p.status = IIC_JPEG_CORRUPT; p.why = "negative file size"; continue; }
    parse_one(blob + offsets[i], size_t(offsets[i + 1] - offsets[i]), p);
    if (p.status == IIC_JPEG_OK) {
      p.img.data_off += uint64_t(offsets[i]);
      pl->ok.push_back(i);
    }
  }

  // Small batches cannot fill the GPU with one serial chain per image (a chain is latency-bound: ~0.15 instructions per clock
  // on a scheduler that could issue 1): split every image into 8 / 4 / 2 chains until the batch alone provides a few thousand of them.
  {
    const char* e = getenv("IIC_JPEG_CHAINS");
    const int want = e ? atoi(e) : 0;
    // ... as long as the per-chain coefficient regions stay within a budget (12 GB unless IIC_JPEG_SCRATCH_GB says otherwise)
    const size_t nimg = pl->ok.size();
    double coef_bytes = 0.0;
    for (size_t j = 0; j < nimg; ++j) {
      const JpegImage& im = pl->files[size_t(pl->ok[j])].img;
      for (int c = 0; c < im.ncomp; ++c) coef_bytes += 128.0 * double(im.comp[c].bw) * double(im.comp[c].bh);
    }
    const char* g = getenv("IIC_JPEG_SCRATCH_GB");
    const double budget = (g && atof(g) > 0.0 ? atof(g) : 12.0) * 1e9;
    int k = 8;
    while (k > 1 && (double(k) * coef_bytes > budget || nimg * size_t(k) > 8192)) k /= 2;
    pl->chains = (want == 1 || want == 2 || want == 4 || want == 8) ? want : k;
    const char* w = getenv("IIC_JPEG_CHAIN_WINDOW");
    if (w && atoi(w) >= 0 && atoi(w) <= 4096) pl->window = atoi(w);
  }
  // the distinct Huffman tables of the batch (most encoders emit the Annex K tables: a handful per batch, L1-resident on the device)
  const size_t m = pl->ok.size();
  {
    std::vector<const HostTable*> seen;
    for (size_t j = 0; j < m; ++j) {
      Parsed& p = pl->files[size_t(pl->ok[j])];
      for (int tc = 0; tc < 2; ++tc)
        for (int th = 0; th < 2; ++th) {
          const HostTable& t = p.tab[tc][th];
          int idx = 0;
          if (t.present) {
            idx = -1;
            for (size_t k = 0; k < seen.size() && idx < 0; ++k)
              if (seen[k]->count == t.count && std::memcmp(seen[k]->bits, t.bits, 16) == 0 &&
                  std::memcmp(seen[k]->vals, t.vals, size_t(t.count)) == 0)
                idx = int(k);
            if (idx < 0) {
              if (seen.size() >= 32000) { delete pl; return IIC_ERR_ARG; }
              idx = int(seen.size());
              seen.push_back(&t);
              pl->tables.emplace_back();
              build_table(t, &pl->tables.back());
              pl->fast.emplace_back();
              if (tc == 1) build_fast_ac(t, &pl->fast.back());
              else std::memset(&pl->fast.back(), 0, sizeof(FastAc));
            }
          }
          p.img.tab[tc * 2 + th] = short(idx);
        }
    }
    if (pl->tables.empty()) { pl->tables.emplace_back(); pl->fast.emplace_back(); }
  }
  // launch order: images with the same AC table pair side by side (they share the CTA's shared-memory copy), the longest
  // entropy-coded segment of a group first
  std::stable_sort(pl->ok.begin(), pl->ok.end(), [&](int a, int b) {
    const JpegImage& x = pl->files[size_t(a)].img;
    const JpegImage& y = pl->files[size_t(b)].img;
    const int kx = (int(x.tab[2]) << 16) | int(x.tab[3]), ky = (int(y.tab[2]) << 16) | int(y.tab[3]);
    return kx != ky ? kx < ky : x.data_len > y.data_len;
  });
  // descriptor region: [JpegImage x m][HuffTable x distinct][order m][units 3m int2][idct_start 3m+1][color_img m][color_start m+1]
  size_t off = 0;
  off += align_up(m * sizeof(JpegImage), 256);
  pl->off_tables = off;
  off += align_up(pl->tables.size() * sizeof(HuffTable), 256);
  pl->off_fast = off;
  off += align_up(pl->fast.size() * sizeof(FastAc), 256);
  pl->off_order = off;
  off += align_up(m * sizeof(int), 256);
  pl->off_units = off;
  off += align_up(3 * m * sizeof(int2), 256);
  pl->off_idct_start = off;
  off += align_up((3 * m + 1) * sizeof(int), 256);
  pl->off_color_img = off;
  off += align_up(m * sizeof(int), 256);
  pl->off_color_start = off;
  off += align_up((m + 1) * sizeof(int), 256);
  pl->desc_bytes = off;
  // coefficient blocks and sample planes behind it
  long long idct_ctas = 0, color_ctas = 0;
  int units = 0;
  for (size_t j = 0; j < m; ++j) {
    Parsed& p = pl->files[size_t(pl->ok[j])];
    size_t all_blocks = 0;
    int bpm = 0;
    for (int c = 0; c < p.img.ncomp; ++c) {
      JpegComp& cp = p.img.comp[c];
      const size_t blocks = size_t(cp.bw) * size_t(cp.bh);
      cp.blk0 = bpm;
      bpm += cp.h * cp.v;
      all_blocks += blocks;
      cp.plane_off = off;
      off += align_up(blocks * 64, 256);
      idct_ctas += (long long)((blocks + 31) / 32);
      ++units;
    }
    const int mcus = p.img.mcux * p.img.mcuy;
    p.img.bpm = bpm;
    // chains inside the image only where each gets a segment several windows long, and never across restart markers
    p.img.nchains = 1;
    if (p.img.restart_interval == 0)
      for (int k = pl->chains; k > 1; k /= 2)
        if (mcus >= k * 4 * kChainWindow) { p.img.nchains = k; break; }
    p.img.info_off = off;
    off += align_up(sizeof(ChainInfo), 256);
    p.img.chain_stride = align_up(all_blocks * 128, 256);
    p.img.coef_off = off;
    off += p.img.chain_stride * size_t(p.img.nchains);
    color_ctas += ((long long)((p.img.width + 3) / 4) * p.img.height + 255) / 256;
  }
  if (idct_ctas > 0x7fffffffLL || color_ctas > 0x7fffffffLL) {
    delete pl;
    return IIC_ERR_ARG;
  }
  pl->scratch_bytes = off + 256;
  pl->idct_ctas = int(idct_ctas);
  pl->color_ctas = int(color_ctas);
  pl->n_units = units;
  *out = pl;
  return IIC_OK;
}

void iic_jpeg_plan_destroy(iic_jpeg_plan* plan) { delete plan; }

int iic_jpeg_plan_info(const iic_jpeg_plan* plan, int i, int* width, int* height, int* status) {
  if (plan == nullptr || i < 0 || size_t(i) >= plan->files.size()) return IIC_ERR_ARG;
  const iic::Parsed& p = plan->files[size_t(i)];
  if (status) *status = p.status;
  if (width) *width = p.status == IIC_JPEG_OK ? p.img.width : 0;
  if (height) *height = p.status == IIC_JPEG_OK ? p.img.height : 0;
  return IIC_OK;
}

int iic_jpeg_plan_infos(const iic_jpeg_plan* plan, int* whs) {
  if (plan == nullptr || whs == nullptr) return IIC_ERR_ARG;
  for (size_t i = 0; i < plan->files.size(); ++i) {
    const iic::Parsed& p = plan->files[i];
    const bool ok = p.status == IIC_JPEG_OK;
    whs[3 * i] = ok ? p.img.width : 0;
    whs[3 * i + 1] = ok ? p.img.height : 0;
    whs[3 * i + 2] = p.status;
  }
  return IIC_OK;
}

const char* iic_jpeg_plan_reason(const iic_jpeg_plan* plan, int i) {
  if (plan == nullptr || i < 0 || size_t(i) >= plan->files.size()) return "";
  return plan->files[size_t(i)].why.c_str();
}

size_t iic_jpeg_plan_staging_bytes(const iic_jpeg_plan* plan) { return plan ? plan->desc_bytes : 0; }
size_t iic_jpeg_plan_scratch_bytes(const iic_jpeg_plan* plan) { return plan ? plan->scratch_bytes : 0; }

int iic_jpeg_decode(const iic_jpeg_plan* plan, const uint8_t* dev_blob, uint8_t* const* out_rgb, void* staging, void* scratch,
                    void* stream) {
  using namespace iic;
  if (plan == nullptr || dev_blob == nullptr || out_rgb == nullptr || staging == nullptr || scratch == nullptr) return IIC_ERR_ARG;
  const size_t m = plan->ok.size();
  if (m == 0) return IIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* st = static_cast<uint8_t*>(staging);
  JpegImage* imgs = reinterpret_cast<JpegImage*>(st);
  HuffTable* tabs = reinterpret_cast<HuffTable*>(st + plan->off_tables);
  int* order = reinterpret_cast<int*>(st + plan->off_order);
  int2* units = reinterpret_cast<int2*>(st + plan->off_units);
  int* idct_start = reinterpret_cast<int*>(st + plan->off_idct_start);
  int* color_img = reinterpret_cast<int*>(st + plan->off_color_img);
  int* color_start = reinterpret_cast<int*>(st + plan->off_color_start);
  int u = 0, ic = 0, cc = 0, m_out = 0;
  for (size_t j = 0; j < m; ++j) {
    const int fi = plan->ok[j];
    const Parsed& p = plan->files[size_t(fi)];
    imgs[j] = p.img;
    imgs[j].out = out_rgb[fi];
    if (out_rgb[fi] == nullptr) continue;     // the caller skips this file
    order[m_out] = int(j);
    for (int c = 0; c < p.img.ncomp; ++c) {
      units[u] = make_int2(int(j), c);
      idct_start[u] = ic;
      ic += int((size_t(p.img.comp[c].bw) * size_t(p.img.comp[c].bh) + 31) / 32);
      ++u;
    }
    color_img[m_out] = int(j);
    color_start[m_out] = cc;
    cc += int(((long long)((p.img.width + 3) / 4) * p.img.height + 255) / 256);
    ++m_out;
  }
  idct_start[u] = ic;
  color_start[m_out] = cc;
  if (m_out == 0) return IIC_OK;
  std::memcpy(tabs, plan->tables.data(), plan->tables.size() * sizeof(HuffTable));
  std::memcpy(st + plan->off_fast, plan->fast.data(), plan->fast.size() * sizeof(FastAc));
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  if (cudaMemcpyAsync(sc, st, plan->desc_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) return IIC_ERR_CUDA;
  const JpegImage* d_imgs = reinterpret_cast<const JpegImage*>(sc);
  {
    const int ipc = kHuffWarps / plan->chains;
    jpeg_huffman_kernel<<<(m_out + ipc - 1) / ipc, 32 * kHuffWarps, 0, s>>>(
        d_imgs, reinterpret_cast<const int*>(sc + plan->off_order), m_out, reinterpret_cast<const HuffTable*>(sc + plan->off_tables),
        reinterpret_cast<const FastAc*>(sc + plan->off_fast), dev_blob, sc, plan->chains, plan->window);
  }
  jpeg_idct_kernel<<<ic, 256, 0, s>>>(d_imgs, reinterpret_cast<const int2*>(sc + plan->off_units),
                                      reinterpret_cast<const int*>(sc + plan->off_idct_start), u, sc);
  jpeg_color_kernel<<<cc, 256, 0, s>>>(d_imgs, reinterpret_cast<const int*>(sc + plan->off_color_img),
                                       reinterpret_cast<const int*>(sc + plan->off_color_start), m_out, sc);
  return cudaGetLastError() == cudaSuccess ? IIC_OK : IIC_ERR_CUDA;
}

}  // extern "C"
