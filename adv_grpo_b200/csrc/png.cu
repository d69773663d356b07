// PNG decode of the reference ("real") images (the reference images of the adversarial loop ARE PNG
// files -- README.md:114-128 of the reference -- opened with `Image.open(fpath).convert("RGB")`,
// scripts/train_sd3_fast_pickscore.py:773-786), hybrid:
//   host  : chunk walk (IHDR / PLTE / IDAT / IEND) and zlib INFLATE of the concatenated IDAT stream (RFC 1950 / 1951: stored,
//           fixed and dynamic Huffman blocks, LZ77 window copies) -- a serial bit-stream problem, plain C++ in this library,
//           no zlib / libpng;
//   device: scan-line UNFILTERING (PNG 1.2 section 6: None / Sub / Up / Average / Paeth).  A reconstructed byte depends on its
//           left, upper and upper-left neighbours, so rows cannot be processed independently; the kernel runs an anti-diagonal
//           WAVEFRONT: thread r owns row r of a band of up to 1024 rows and at step t reconstructs pixel t - r; the pixel above
//           arrives from the neighbouring thread through shared memory (double-buffered by step parity), left / upper-left
//           stay in registers; W + rows - 1 steps of one barrier each.  Then conversion to interleaved RGB bytes (truecolour:
//           in place; alpha dropped; greyscale replicated; palette looked up), as Pillow's convert("RGB") does.
// Every colour type and bit depth of PNG 1.2 (1 / 2 / 4 / 8 / 16-bit greyscale, 8 / 16-bit greyscale + alpha, 8 / 16-bit
// truecolour (+ alpha), 1..8-bit palette), non-interlaced or Adam7-interlaced (seven reduced images, each unfiltered on its
// own and scattered into place).
// The file bytes are untrusted: the host half checks every chunk checksum, the chunk order, zlib's code-completeness rules, the
// Adler-32 trailer and the filter types, and is fuzzed under AddressSanitizer and differentially against Pillow
// (tests/fuzz/fuzz_image_decoders.cpp, tests/test_decoder_fuzz.py).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace advgrpo {
namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// CRC-32 of a chunk's type + data (PNG 1.2 section 3.4), eight bytes per step
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (int s = 1; s < 8; ++s)
      for (uint32_t i = 0; i < 256; ++i) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 255];
  }
};

uint32_t crc32_of(const uint8_t* p, size_t n) {
  static const CrcTables T;
  uint32_t c = 0xFFFFFFFFu;
  for (; n >= 8; n -= 8, p += 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T.t[7][lo & 255] ^ T.t[6][(lo >> 8) & 255] ^ T.t[5][(lo >> 16) & 255] ^ T.t[4][lo >> 24] ^ T.t[3][hi & 255] ^
        T.t[2][(hi >> 8) & 255] ^ T.t[1][(hi >> 16) & 255] ^ T.t[0][hi >> 24];
  }
  for (; n; --n, ++p) c = T.t[0][(c ^ *p) & 255] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

// Adler-32 of the inflated data (RFC 1950), modulo deferred over runs of 5552 bytes
uint32_t adler32_of(const uint8_t* p, size_t n) {
  uint32_t a = 1, b = 0;
  while (n) {
    size_t k = n < 5552 ? n : 5552;
    n -= k;
    for (; k >= 8; k -= 8, p += 8) {
      a += p[0]; b += a; a += p[1]; b += a; a += p[2]; b += a; a += p[3]; b += a;
      a += p[4]; b += a; a += p[5]; b += a; a += p[6]; b += a; a += p[7]; b += a;
    }
    for (; k; --k) { a += *p++; b += a; }
    a %= 65521u;
    b %= 65521u;
  }
  return (b << 16) | a;
}

int png_channels(int ct) { return ct == 0 ? 1 : ct == 2 ? 3 : ct == 3 ? 1 : ct == 4 ? 2 : ct == 6 ? 4 : 0; }

// Adam7 (PNG 1.2 section 8.2): pass p covers pixels (x0 + i dx, y0 + j dy)
const int kAdam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};

struct PngPass { int x0, y0, dx, dy, w, h; int64_t rowbytes; };

// the reduced images of an interlaced file (or the one full image of a non-interlaced one) that are not empty
int png_passes(const advgrpo_png_info& I, PngPass* out) {
  const int bits = I.channels * I.bit_depth;
  if (!I.interlace) {
    out[0] = {0, 0, 1, 1, I.width, I.height, (int64_t)I.rowbytes};
    return 1;
  }
  int n = 0;
  for (int p = 0; p < 7; ++p) {
    const int x0 = kAdam7[p][0], y0 = kAdam7[p][1], dx = kAdam7[p][2], dy = kAdam7[p][3];
    const int w = (I.width - x0 + dx - 1) / dx, h = (I.height - y0 + dy - 1) / dy;
    if (w <= 0 || h <= 0) continue;
    out[n++] = {x0, y0, dx, dy, w, h, ((int64_t)w * bits + 7) / 8};
  }
  return n;
}

struct PngParsed {
  advgrpo_png_info info;
  std::vector<std::pair<size_t, size_t>> idat;   // (offset, length) of every IDAT body
  uint8_t palette[768];
};

// 0 ok, negative error; info.supported says whether this decoder takes the file
int parse_png(const uint8_t* d, size_t n, PngParsed& P) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  memset(&P.info, 0, sizeof(P.info));
  memset(P.palette, 0, sizeof(P.palette));
  P.idat.clear();
  if (n < 8 || memcmp(d, sig, 8)) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: not a PNG file");
  size_t pos = 8;
  bool have_ihdr = false, have_plte = false, have_iend = false, idat_closed = false;
  // The chunk walk is stricter than Pillow's (which skips the IDAT and trailing CRCs): whatever it refuses goes to the caller's
  // host decoder, so a file this decoder takes is one every decoder agrees on.
  while (pos + 12 <= n) {
    const size_t ln = be32(d + pos);
    const uint8_t* typ = d + pos + 4;
    if (ln > 0x7FFFFFFFu || pos + 12 + ln > n) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: truncated chunk");
    for (int k = 0; k < 4; ++k)
      if (!((typ[k] >= 'A' && typ[k] <= 'Z') || (typ[k] >= 'a' && typ[k] <= 'z'))) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad chunk type");
    if (crc32_of(typ, 4 + ln) != be32(d + pos + 8 + ln)) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: chunk checksum mismatch (%.4s)", (const char*)typ);
    const uint8_t* body = d + pos + 8;
    const bool is_idat = !memcmp(typ, "IDAT", 4);
    if (!have_ihdr && memcmp(typ, "IHDR", 4)) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: the first chunk is not IHDR");
    if (!P.idat.empty() && !is_idat) idat_closed = true;
    if (!memcmp(typ, "IHDR", 4)) {
      if (ln != 13 || have_ihdr) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad IHDR");
      P.info.width = (int32_t)be32(body);
      P.info.height = (int32_t)be32(body + 4);
      P.info.bit_depth = body[8];
      P.info.color_type = body[9];
      if (body[10] != 0 || body[11] != 0) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: unknown compression / filter method");
      P.info.interlace = body[12];
      if (be32(body) > (1u << 15) || be32(body + 4) > (1u << 15)) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad image size");
      have_ihdr = true;
    } else if (!memcmp(typ, "PLTE", 4)) {
      if (have_plte || !P.idat.empty() || ln == 0 || ln > 768 || ln % 3) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad PLTE");
      memcpy(P.palette, body, ln);
      P.info.palette_entries = (int32_t)(ln / 3);
      have_plte = true;
    } else if (is_idat) {
      if (idat_closed) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: IDAT chunks are not consecutive");
      P.idat.push_back({pos + 8, ln});
    } else if (!memcmp(typ, "IEND", 4)) {
      have_iend = true;
      break;
    } else if (!memcmp(typ, "acTL", 4)) {
      P.info.supported = 0;                                   // animated PNG: Pillow's frame logic decides
      return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: animated PNG");
    }
    pos += 12 + ln;
  }
  if (!have_iend) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: no IEND chunk (truncated file)");
  if (!have_ihdr || P.idat.empty()) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: missing IHDR or IDAT");
  if (P.info.width < 1 || P.info.height < 1 || P.info.width > (1 << 15) || P.info.height > (1 << 15))
    return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad image size");
  if ((int64_t)P.info.width * P.info.height > 2 * (int64_t)89478485)      // Pillow's DecompressionBombError bound
    return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: image larger than 178956970 pixels");
  P.info.channels = png_channels(P.info.color_type);
  if (!P.info.channels) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad colour type");
  const int bd = P.info.bit_depth, ct = P.info.color_type;
  const bool depth_ok = (ct == 0 && (bd == 1 || bd == 2 || bd == 4 || bd == 8 || bd == 16)) || ((ct == 2 || ct == 6) && (bd == 8 || bd == 16)) ||
                        (ct == 3 && (bd == 1 || bd == 2 || bd == 4 || bd == 8)) || (ct == 4 && (bd == 8 || bd == 16));
  if (bd != 1 && bd != 2 && bd != 4 && bd != 8 && bd != 16) return set_error(ADVGRPO_ERR_BAD_ARG, "png_parse: bad bit depth");
  P.info.rowbytes = (int32_t)(((int64_t)P.info.width * P.info.channels * bd + 7) / 8);
  P.info.supported = depth_ok && (P.info.interlace == 0 || P.info.interlace == 1) && (ct != 3 || have_plte);
  return ADVGRPO_OK;
}

// ---- inflate (RFC 1951), LSB-first bit reader over the concatenated IDAT bodies -------------------------------------------
struct InBits {
  const uint8_t* d;
  const std::vector<std::pair<size_t, size_t>>* segs;
  size_t seg, off;
  uint64_t acc;
  int cnt;
  bool eof;
  int fake = 0;                                 // zero bits fed behind the end of the data (the top of acc): cnt < fake = some were used
  int next_byte() {
    while (seg < segs->size() && off >= (*segs)[seg].second) { ++seg; off = 0; }
    if (seg >= segs->size()) { eof = true; fake += 8; return 0; }
    return d[(*segs)[seg].first + off++];
  }
  inline void need(int k) {
    if (cnt >= k) return;
    if (seg < segs->size() && off + 8 <= (*segs)[seg].second) {      // eight bytes of the current IDAT body at once
      uint64_t v;
      memcpy(&v, d + (*segs)[seg].first + off, 8);                 // little-endian host (x86-64 / aarch64)
      const int adv = (63 - cnt) >> 3;
      acc |= v << cnt;
      cnt += 8 * adv;
      off += adv;
      acc &= (1ull << cnt) - 1;                                    // cnt <= 63: drop the bits of bytes not consumed yet
      return;
    }
    while (cnt < k) { acc |= (uint64_t)next_byte() << cnt; cnt += 8; }
  }
  inline uint32_t bits(int k) {
    if (!k) return 0;
    need(k);
    const uint32_t v = (uint32_t)(acc & ((1ull << k) - 1));
    acc >>= k;
    cnt -= k;
    return v;
  }
  inline void align_byte() { const int r = cnt & 7; acc >>= r; cnt -= r; }
};

struct HuffDec {
  uint16_t count[16], symbol[288];
  uint16_t fast[1024];      // index = next 10 bits (LSB first); (length << 9) | symbol, 0 = longer / invalid
  // complete_or_single: zlib's rule (inftrees.c) -- an incomplete code is an error unless it is a literal / distance code
  // made of one single 1-bit code; false for the fixed distance code (30 of 32 codes)
  bool build(const uint8_t* lengths, int n, bool is_code_length_code = false, bool check_complete = true) {
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; ++i) count[lengths[i]]++;
    if (count[0] == n) { memset(fast, 0, sizeof(fast)); return true; }          // no codes: decoding one reports the error
    int left = 1, max_len = 0;
    for (int l = 1; l < 16; ++l) {
      left <<= 1;
      left -= count[l];
      if (left < 0) return false;                                              // over-subscribed
      if (count[l]) max_len = l;
    }
    if (check_complete && left > 0 && (is_code_length_code || max_len != 1)) return false;
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (int i = 0; i < n; ++i)
      if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    // fast table: canonical codes, bit-reversed (deflate packs Huffman codes starting from their most significant bit)
    memset(fast, 0, sizeof(fast));
    int code = 0, idx = 0;
    for (int l = 1; l <= 10; ++l) {
      for (int i = 0; i < count[l]; ++i, ++code, ++idx) {
        int rev = 0;
        for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1) << (l - 1 - b);
        for (int f = rev; f < 1024; f += 1 << l) fast[f] = (uint16_t)((l << 9) | symbol[idx]);
      }
      code <<= 1;
    }
    return true;
  }
  inline int decode(InBits& br) const {
    br.need(15);
    const uint16_t f = fast[br.acc & 1023];
    if (f) { br.acc >>= (f >> 9); br.cnt -= (f >> 9); return f & 511; }
    int code = 0, first = 0, index = 0;                                        // canonical bit-by-bit (codes of 11..15 bits)
    for (int l = 1; l < 16; ++l) {
      code |= (int)(br.acc & 1);
      br.acc >>= 1;
      br.cnt -= 1;
      const int c = count[l];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};

int inflate_stream(InBits& br, uint8_t* out, size_t out_cap, size_t* out_len) {
  static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  HuffDec* lit = new HuffDec();
  HuffDec* dist = new HuffDec();
  size_t o = 0;
  int rc = ADVGRPO_OK;
  auto fail = [&](const char* why) { rc = set_error(ADVGRPO_ERR_BAD_ARG, "png_inflate: %s", why); };
  for (bool last = false; !last && rc == ADVGRPO_OK;) {
    last = br.bits(1);
    const int type = (int)br.bits(2);
    if (type == 0) {
      br.align_byte();
      const uint32_t len = br.bits(16), nlen = br.bits(16);
      if ((len ^ 0xFFFF) != nlen) { fail("stored block length check"); break; }
      if (o + len > out_cap) { fail("output overflow"); break; }
      for (uint32_t i = 0; i < len; ++i) out[o++] = (uint8_t)br.bits(8);
      continue;
    }
    if (type == 3) { fail("reserved block type"); break; }
    uint8_t lengths[320];
    if (type == 1) {
      for (int i = 0; i < 144; ++i) lengths[i] = 8;
      for (int i = 144; i < 256; ++i) lengths[i] = 9;
      for (int i = 256; i < 280; ++i) lengths[i] = 7;
      for (int i = 280; i < 288; ++i) lengths[i] = 8;
      lit->build(lengths, 288);
      for (int i = 0; i < 30; ++i) lengths[i] = 5;
      dist->build(lengths, 30, false, false);
    } else {
      const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
      if (nlen > 286 || ndist > 30) { fail("bad table sizes"); break; }
      uint8_t cl[19];
      memset(cl, 0, sizeof(cl));
      for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
      HuffDec* lencode = new HuffDec();
      if (!lencode->build(cl, 19, true)) { delete lencode; fail("bad code-length code"); break; }
      int idx = 0;
      while (idx < nlen + ndist) {
        int sym = lencode->decode(br);
        if (sym < 0) { fail("bad code-length symbol"); break; }
        if (sym < 16) { lengths[idx++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
          if (idx == 0) { fail("repeat without a previous length"); break; }
          val = lengths[idx - 1];
          rep = 3 + (int)br.bits(2);
        } else if (sym == 17) rep = 3 + (int)br.bits(3);
        else rep = 11 + (int)br.bits(7);
        if (idx + rep > nlen + ndist) { fail("too many lengths"); break; }
        while (rep--) lengths[idx++] = (uint8_t)val;
      }
      delete lencode;
      if (rc != ADVGRPO_OK) break;
      if (lengths[256] == 0) { fail("no end-of-block code"); break; }
      if (!lit->build(lengths, nlen) || !dist->build(lengths + nlen, ndist)) { fail("bad Huffman table"); break; }
    }
    for (;;) {
      const int sym = lit->decode(br);
      if (sym < 0) { fail("bad literal / length code"); break; }
      if (sym < 256) {
        if (o >= out_cap) { fail("output overflow"); break; }
        out[o++] = (uint8_t)sym;
      } else if (sym == 256) {
        break;
      } else {
        const int ls = sym - 257;
        if (ls >= 29) { fail("bad length symbol"); break; }
        const int len = lbase[ls] + (int)br.bits(lext[ls]);
        const int ds = dist->decode(br);
        if (ds < 0 || ds >= 30) { fail("bad distance code"); break; }
        const size_t dd = dbase[ds] + br.bits(dext[ds]);
        if (dd > o) { fail("distance beyond the start of the output"); break; }
        if (o + len > out_cap) { fail("output overflow"); break; }
        if (dd >= 8 && o + len + 8 <= out_cap) {                     // eight bytes per step; may run up to 7 bytes past o + len
          const uint8_t* src = out + o - dd;
          uint8_t* dst = out + o;
          for (int i = 0; i < len; i += 8) memcpy(dst + i, src + i, 8);
          o += len;
        } else {
          for (int i = 0; i < len; ++i, ++o) out[o] = out[o - dd];
        }
      }
    }
  }
  delete lit;
  delete dist;
  *out_len = o;
  return rc;
}

// ---- device: wavefront unfilter + RGB conversion ---------------------------------------------------------------------------
__device__ __forceinline__ int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// raw: [H, 1 + rowbytes] filtered scan lines; rows: [H, rowbytes] reconstructed bytes.  ONE block; thread r = row band0 + r.
// BPP = bytes per complete pixel as the filters see it (1 for sub-byte depths, 6 / 8 for 16-bit truecolour); W = rowbytes / BPP.
template <int BPP>
__global__ void __launch_bounds__(1024)
png_unfilter_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ rows, int W, int H) {
  __shared__ uint64_t up_sm[2][1024];                       // the pixel each row reconstructed in the previous step, by parity
  const int r = threadIdx.x, R = blockDim.x;
  const int64_t rowbytes = (int64_t)W * BPP;
  for (int band0 = 0; band0 < H; band0 += R) {
    const int y = band0 + r;
    const bool live = y < H;
    const int rows_here = min(R, H - band0);
    const uint8_t* in = raw + (int64_t)(live ? y : 0) * (1 + rowbytes);
    uint8_t* out = rows + (int64_t)(live ? y : 0) * rowbytes;
    const uint8_t* prior = rows + (int64_t)(y - 1) * rowbytes;      // only read by the first row of a band (y > 0)
    const int ft = live ? in[0] : 0;
    uint64_t left = 0, upleft = 0;                            // packed BPP bytes
    const int steps = W + rows_here - 1;
    // The filtered bytes of a pixel are fetched four steps before the wavefront reaches it (a queue of four packed pixels in
    // registers), so the global-memory latency of the row-strided loads is off the step's critical path.
    auto fetch = [&](int p) -> uint64_t {
      uint64_t v = 0;
      if (live && p >= 0 && p < W) {
#pragma unroll
        for (int k = 0; k < BPP; ++k) v |= (uint64_t)in[1 + (int64_t)p * BPP + k] << (8 * k);
      }
      return v;
    };
    uint64_t q0 = fetch(-r), q1 = fetch(1 - r), q2 = fetch(2 - r), q3 = fetch(3 - r);
    for (int t = 0; t < steps; ++t) {
      const int px = t - r;
      const uint64_t cur = q0;                                // the filtered bytes of pixel px
      q0 = q1;
      q1 = q2;
      q2 = q3;
      q3 = fetch(px + 4);
      uint64_t rec = 0;
      if (live && px >= 0 && px < W) {
        uint64_t up = 0;
        if (r > 0) up = up_sm[(t + 1) & 1][r - 1];            // written by row r - 1 in step t - 1
        else if (y > 0) {
#pragma unroll
          for (int k = 0; k < BPP; ++k) up |= (uint64_t)prior[(int64_t)px * BPP + k] << (8 * k);
        }
#pragma unroll
        for (int k = 0; k < BPP; ++k) {
          const int x = (int)((cur >> (8 * k)) & 255);
          const int a = (int)((left >> (8 * k)) & 255), b = (int)((up >> (8 * k)) & 255), c = (int)((upleft >> (8 * k)) & 255);
          int pr = 0;
          if (ft == 1) pr = a;
          else if (ft == 2) pr = b;
          else if (ft == 3) pr = (a + b) >> 1;
          else if (ft == 4) pr = paeth(a, b, c);
          const int v = (x + pr) & 255;
          rec |= (uint64_t)v << (8 * k);
          out[(int64_t)px * BPP + k] = (uint8_t)v;
        }
        left = rec;
        upleft = up;
      }
      up_sm[t & 1][r] = rec;
      __syncthreads();
    }
    __syncthreads();                                          // the band's last row is complete in global memory
  }
}

// rows [H, rowbytes] -> rgb [H, W, 3], as Pillow's convert("RGB"): sub-byte samples unpacked MSB first (greyscale scaled to
// 0..255, palette indices looked up), 16-bit truecolour keeps the high byte of every sample, 16-bit greyscale is CLIPPED to 255
// (Pillow's I;16 -> RGB), alpha dropped, grey replicated.
__global__ void __launch_bounds__(256)
png_to_rgb_kernel(const uint8_t* __restrict__ rows, const uint8_t* __restrict__ palette, uint8_t* __restrict__ rgb, int W, int H,
                  int64_t rowbytes, int ch, int color_type, int bd, int x0, int y0, int dx, int dy, int out_w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)W * H) return;
  const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
  const int64_t o = ((int64_t)(y0 + y * dy) * out_w + (x0 + x * dx)) * 3;      // Adam7: scatter the reduced image into place
  const uint8_t* row = rows + (int64_t)y * rowbytes;
  int v[3];
  if (bd < 8) {
    const int bit = x * bd;
    const int s = (row[bit >> 3] >> (8 - bd - (bit & 7))) & ((1 << bd) - 1);
    if (color_type == 3) { v[0] = palette[3 * s]; v[1] = palette[3 * s + 1]; v[2] = palette[3 * s + 2]; }
    else v[0] = v[1] = v[2] = s * (255 / ((1 << bd) - 1));
  } else {
    const int sb = bd >> 3;                                   // bytes per sample
    const uint8_t* p = row + (int64_t)x * ch * sb;
    if (color_type == 2 || color_type == 6) { v[0] = p[0]; v[1] = p[sb]; v[2] = p[2 * sb]; }
    else if (color_type == 3) { v[0] = palette[3 * p[0]]; v[1] = palette[3 * p[0] + 1]; v[2] = palette[3 * p[0] + 2]; }
    else {
      const int g = (sb == 2 && color_type == 0) ? min(p[0] * 256 + p[1], 255) : p[0];   // I;16 is clipped, LA;16B keeps the high byte
      v[0] = v[1] = v[2] = g;
    }
  }
  rgb[o] = (uint8_t)v[0];
  rgb[o + 1] = (uint8_t)v[1];
  rgb[o + 2] = (uint8_t)v[2];
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

int advgrpo_png_parse(const uint8_t* file, size_t nbytes, advgrpo_png_info* info) {
  ADVGRPO_CHECK_ARG(file && info, "png_parse: null pointer");
  PngParsed* P = new PngParsed();
  const int rc = parse_png(file, nbytes, *P);
  *info = P->info;
  delete P;
  return rc;
}

size_t advgrpo_png_raw_bytes(const advgrpo_png_info* info) {
  if (!info || !info->supported) return 0;
  PngPass ps[7];
  const int np = png_passes(*info, ps);
  size_t n = 0;
  for (int p = 0; p < np; ++p) n += (size_t)ps[p].h * (1 + (size_t)ps[p].rowbytes);
  return n;
}

int advgrpo_png_inflate(const uint8_t* file, size_t nbytes, uint8_t* raw_host, uint8_t* palette_host) {
  ADVGRPO_CHECK_ARG(file && raw_host && palette_host, "png_inflate: null pointer");
  PngParsed* P = new PngParsed();
  int rc = parse_png(file, nbytes, *P);
  if (rc == ADVGRPO_OK && !P->info.supported) rc = set_error(ADVGRPO_ERR_UNSUPPORTED, "png_inflate: file is outside the supported subset");
  if (rc != ADVGRPO_OK) { delete P; return rc; }
  memcpy(palette_host, P->palette, 768);
  InBits br{file, &P->idat, 0, 0, 0, 0, false};
  const uint32_t cmf = br.bits(8), flg = br.bits(8);                       // zlib header (RFC 1950)
  if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 32)) {
    delete P;
    return set_error(ADVGRPO_ERR_BAD_ARG, "png_inflate: bad zlib header");
  }
  const size_t want = advgrpo_png_raw_bytes(&P->info);
  size_t got = 0;
  rc = inflate_stream(br, raw_host, want, &got);
  if (rc == ADVGRPO_OK && got != want) rc = set_error(ADVGRPO_ERR_BAD_ARG, "png_inflate: %zu bytes of image data, %zu expected", got, want);
  if (rc == ADVGRPO_OK) {                                      // zlib trailer: Adler-32 of the inflated data, big-endian
    br.align_byte();
    uint32_t sum = 0;
    for (int k = 0; k < 4; ++k) sum = (sum << 8) | br.bits(8);
    if (br.cnt < br.fake || sum != adler32_of(raw_host, got)) rc = set_error(ADVGRPO_ERR_BAD_ARG, "png_inflate: Adler-32 mismatch");
  }
  if (rc == ADVGRPO_OK) {                                      // every scan line starts with a filter type 0..4 (PNG 1.2 section 6.1)
    PngPass ps[7];
    const int np = png_passes(P->info, ps);
    const uint8_t* line = raw_host;
    for (int q = 0; q < np && rc == ADVGRPO_OK; ++q)
      for (int y = 0; y < ps[q].h; ++y, line += 1 + ps[q].rowbytes)
        if (*line > 4) { rc = set_error(ADVGRPO_ERR_BAD_ARG, "png_inflate: filter type %d", (int)*line); break; }
  }
  delete P;
  return rc;
}

size_t advgrpo_png_workspace_bytes(const advgrpo_png_info* info) {
  if (!info || !info->supported) return 0;
  return advgrpo_png_raw_bytes(info) + 256;                  // the reconstructed scan lines of every pass (upper bound)
}

int advgrpo_png_unfilter_to_rgb(const uint8_t* raw_dev, const uint8_t* palette_dev, const advgrpo_png_info* info,
                                uint8_t* rgb_hwc_dev, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(raw_dev && info && rgb_hwc_dev, "png_unfilter_to_rgb: null pointer");
  ADVGRPO_CHECK_ARG(info->supported && info->width >= 1 && info->height >= 1 && info->channels >= 1 && info->channels <= 4 &&
                        info->rowbytes >= 1,
                    "png_unfilter_to_rgb: unsupported file (advgrpo_png_parse reported supported = 0)");
  ADVGRPO_CHECK_ARG(info->color_type != 3 || palette_dev, "png_unfilter_to_rgb: palette image without a palette");
  cudaStream_t st = (cudaStream_t)stream;
  const bool direct = info->color_type == 2 && info->bit_depth == 8 && !info->interlace;   // rows ARE the RGB bytes
  if (!direct && (!workspace || workspace_bytes < advgrpo_png_workspace_bytes(info)))
    return set_error(ADVGRPO_ERR_WORKSPACE, "png_unfilter_to_rgb: workspace too small");
  const int bits = info->channels * info->bit_depth;
  const int fbpp = bits >= 8 ? bits / 8 : 1;                   // bytes per complete pixel (PNG 1.2 section 6.2), at least 1
  PngPass ps[7];
  const int np = png_passes(*info, ps);
  size_t raw_off = 0, row_off = 0;
  for (int p = 0; p < np; ++p) {
    const PngPass& q = ps[p];
    const uint8_t* raw_p = raw_dev + raw_off;
    uint8_t* rows = direct ? rgb_hwc_dev : (uint8_t*)workspace + row_off;
    const int threads = q.h < 1024 ? ((q.h + 31) / 32) * 32 : 1024;
    const int units = (int)(q.rowbytes / fbpp);
    switch (fbpp) {
      case 1: png_unfilter_kernel<1><<<1, threads, 0, st>>>(raw_p, rows, units, q.h); break;
      case 2: png_unfilter_kernel<2><<<1, threads, 0, st>>>(raw_p, rows, units, q.h); break;
      case 3: png_unfilter_kernel<3><<<1, threads, 0, st>>>(raw_p, rows, units, q.h); break;
      case 4: png_unfilter_kernel<4><<<1, threads, 0, st>>>(raw_p, rows, units, q.h); break;
      case 6: png_unfilter_kernel<6><<<1, threads, 0, st>>>(raw_p, rows, units, q.h); break;
      case 8: png_unfilter_kernel<8><<<1, threads, 0, st>>>(raw_p, rows, units, q.h); break;
      default: return set_error(ADVGRPO_ERR_UNSUPPORTED, "png_unfilter_to_rgb: %d bytes per pixel", fbpp);
    }
    ADVGRPO_CUDA_LAUNCH_CHECK();
    if (!direct) {
      const int64_t npx = (int64_t)q.w * q.h;
      png_to_rgb_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(rows, palette_dev, rgb_hwc_dev, q.w, q.h, q.rowbytes,
                                                                      info->channels, info->color_type, info->bit_depth, q.x0,
                                                                      q.y0, q.dx, q.dy, info->width);
      ADVGRPO_CUDA_LAUNCH_CHECK();
    }
    raw_off += (size_t)q.h * (1 + (size_t)q.rowbytes);
    row_off += (size_t)q.h * (size_t)q.rowbytes;
  }
  return ADVGRPO_OK;
}

}  // extern "C"
