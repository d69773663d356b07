// Baseline JPEG decode of the reference ("real") images (the reference does
// `Image.open(fpath).convert("RGB")` per file per batch, scripts/train_sd3_fast_pickscore.py:773-786), hybrid like every
// production GPU decoder:
//   host  : marker parse + sequential Huffman entropy decode (ITU T.81 F.2.2; a serial bit-stream problem) into int16
//           coefficient blocks -- plain C++ in this library, no libjpeg;
//   device: dequantisation + the "islow" integer inverse DCT (8x8 blocks), "fancy" triangle chroma upsampling
//           (4:2:0 / 4:2:2), fixed-point YCbCr -> RGB, interleaved uint8 output -- integer work, one thread per block /
//           per pixel, coalesced byte stores.
// Bit-exact with libjpeg(-turbo)'s default decompression settings (JDCT_ISLOW, do_fancy_upsampling), i.e. with what Pillow
// returns for the same file: jidctint.c (CONST_BITS 13, PASS1_BITS 2), jdsample.c h2v1 / h2v2_fancy_upsample, jdcolor.c
// build_ycc_rgb_table.  Sequential (SOF0 / SOF1) and PROGRESSIVE (SOF2: DC / AC first and refinement scans with end-of-band
// runs, jdphuff.c) Huffman files; files outside the supported subset (arithmetic, lossless, 12-bit, CMYK / RGB colour spaces,
// non-interleaved sequential scans, chroma factors other than 1x1) are reported as unsupported by advgrpo_jpeg_parse and stay
// on the caller's host decoder.
// The file bytes are untrusted.  The host half never reads or writes outside its buffers (fuzzed under AddressSanitizer,
// tests/fuzz/fuzz_image_decoders.cpp) and takes only streams that every decoder reads the same way: a scan segment that does
// not end on its last block, a restart marker out of sequence, a missing EOI, a Huffman table libjpeg refuses, an unknown
// marker, an incomplete progressive script or coefficients beyond the range of 8-bit samples is an error, so that whatever
// this decoder accepts comes out as Pillow decodes it (tests/test_decoder_fuzz.py).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace advgrpo {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Pillow refuses files beyond 2 x Image.MAX_IMAGE_PIXELS (DecompressionBombError); the host decoders use the same bound
const int64_t kMaxPixels = 2 * (int64_t)89478485;
// Dequantised coefficients of 8-bit samples have an L2 norm of at most 1024 per block, so the absolute sum of a column stays
// below 2897 (+ quantisation error).  Beyond 4096 the file is corrupt, and libjpeg-turbo's 16-bit SIMD inverse DCT (which Pillow
// runs) starts to wrap / saturate where the 32-bit arithmetic of the device kernel does not: such files go to the host decoder.
const int kMaxColumnSum = 4096;

struct HuffTable {
  bool present = false;
  uint8_t counts[16];
  uint8_t symbols[256];
  // canonical decoding (T.81 F.2.2.3): per length the smallest / largest code and the index of its first symbol
  int32_t mincode[17], maxcode[18], valptr[17];
  // 9-bit lookahead: (length << 8) | symbol, 0 = longer code
  uint16_t fast[512];
  // AC tables, 10-bit lookahead over code + magnitude bits together: (value << 16) | (run << 8) | total bits, 0 = take the long way
  int32_t fast_ac[1024];
};

struct Scan {
  int ns;                      // components in the scan
  int ci[3], td[3], ta[3];     // component index (frame order) and DC / AC table selectors, per scan component
  int ss, se, ah, al;          // spectral selection and successive approximation (0, 63, 0, 0 for sequential files)
  int ri;                      // restart interval in force
  size_t ecs;                  // offset of the entropy-coded segment
  int dc_idx[4], ac_idx[4];    // Huffman tables in force (indices into Parsed::pool; -1 = undefined)
};

struct Parsed {
  advgrpo_jpeg_info info;
  uint16_t qt[4][64];          // natural order
  bool qt_present[4];
  std::vector<HuffTable> pool; // every DHT definition met (tables may be redefined between scans)
  int cur_dc[4], cur_ac[4];
  std::vector<Scan> scans;
};

// false: the code lengths over-subscribe the code space (not a prefix code) -- a corrupt DHT segment
bool build_huff(HuffTable& t) {
  int code = 0, k = 0, lmax = 0;
  for (int l = 1; l <= 16; ++l)
    if (t.counts[l - 1]) lmax = l;
  for (int l = 1; l <= 16; ++l) {
    t.valptr[l] = k;
    t.mincode[l] = code;
    code += t.counts[l - 1];
    k += t.counts[l - 1];
    if (l <= lmax && code >= (1 << l)) return false;     // jdhuff.c jpeg_make_d_derived_tbl: the all-ones code is not allowed either
    t.maxcode[l] = t.counts[l - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
  memset(t.fast, 0, sizeof(t.fast));
  code = 0;
  k = 0;
  for (int l = 1; l <= 9; ++l) {
    for (int i = 0; i < t.counts[l - 1]; ++i, ++k, ++code) {
      const int base = code << (9 - l);
      for (int f = 0; f < (1 << (9 - l)); ++f) t.fast[base + f] = (uint16_t)((l << 8) | t.symbols[k]);
    }
    code <<= 1;
  }
  for (int i = 0; i < 1024; ++i) {
    t.fast_ac[i] = 0;
    const uint16_t f = t.fast[i >> 1];
    const int len = f >> 8, run = (f >> 4) & 15, mag = f & 15;
    if (!f || !mag || len + mag > 10) continue;
    int v = ((i << len) & 1023) >> (10 - mag);
    if (v < (1 << (mag - 1))) v -= (1 << mag) - 1;
    t.fast_ac[i] = (int32_t)((uint32_t)v << 16) | (run << 8) | (len + mag);
  }
  t.present = true;
  return true;
}

// returns 0 ok, 1 unsupported (info.supported = 0), negative error
int parse_jpeg(const uint8_t* d, size_t n, Parsed& P) {
  memset(&P.info, 0, sizeof(P.info));
  memset(P.qt_present, 0, sizeof(P.qt_present));
  for (int i = 0; i < 4; ++i) P.cur_dc[i] = P.cur_ac[i] = -1;
  P.pool.clear();
  P.scans.clear();
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: not a JPEG file (no SOI)");
  size_t pos = 2;
  bool have_frame = false, saw_jfif = false, saw_adobe = false;
  int adobe_transform = -1, comp_id[3] = {0, 0, 0}, ri = 0;
  int coef_bits[3][64];
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 64; ++k) coef_bits[c][k] = -1;
  auto unsupported = [&](const char* why) {
    P.info.supported = 0;
    set_error(ADVGRPO_ERR_UNSUPPORTED, "jpeg_parse: %s", why);
    return 1;
  };
  while (pos + 2 <= n) {
    if (d[pos] != 0xFF) {                        // entropy-coded bytes of the previous scan: on to the next marker
      if (P.scans.empty()) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: marker expected at byte %zu", pos);
      ++pos;
      continue;
    }
    const uint8_t m = d[pos + 1];
    if (m == 0xFF) { ++pos; continue; }
    if (m == 0x00 || (m >= 0xD0 && m <= 0xD7)) { pos += 2; continue; }   // stuffed byte / RSTn inside a scan
    pos += 2;
    if (m == 0x01) continue;
    if (m == 0xD9) break;                        // EOI
    if (pos + 2 > n) break;
    const size_t ln = ((size_t)d[pos] << 8) | d[pos + 1];
    if (ln < 2 || pos + ln > n) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: truncated segment");
    const uint8_t* s = d + pos + 2;
    const size_t sl = ln - 2;
    if (m == 0xDB) {
      if (!P.scans.empty()) return unsupported("quantisation table redefined between scans");
      size_t i = 0;
      while (i + 65 <= sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        if (pq) return unsupported("16-bit quantisation table");
        if (tq > 3) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad DQT id");
        for (int k = 0; k < 64; ++k) P.qt[tq][kZigzag[k]] = s[i + 1 + k];
        P.qt_present[tq] = true;
        i += 65;
      }
      if (i != sl) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad DQT length");
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
      if (sl < 6) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: short SOF");
      if (have_frame) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: more than one SOF marker");
      if (s[0] != 8) return unsupported("sample precision other than 8 bits");
      P.info.progressive = m == 0xC2;
      P.info.height = (s[1] << 8) | s[2];
      P.info.width = (s[3] << 8) | s[4];
      P.info.ncomp = s[5];
      if (P.info.ncomp != 1 && P.info.ncomp != 3) return unsupported("component count other than 1 or 3 (CMYK / YCCK)");
      if (sl != (size_t)(6 + 3 * P.info.ncomp) || P.info.width < 1 || P.info.height < 1)
        return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad SOF");
      if ((int64_t)P.info.width * P.info.height > kMaxPixels)
        return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: image larger than %lld pixels", (long long)kMaxPixels);
      for (int c = 0; c < P.info.ncomp; ++c) {
        comp_id[c] = s[6 + 3 * c];
        P.info.h[c] = s[7 + 3 * c] >> 4;
        P.info.v[c] = s[7 + 3 * c] & 15;
        P.info.tq[c] = s[8 + 3 * c];
        if (P.info.tq[c] > 3) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad quantisation table id");
        if (P.info.h[c] < 1 || P.info.h[c] > 4 || P.info.v[c] < 1 || P.info.v[c] > 4)      // jdinput.c: JERR_BAD_SAMPLING
          return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: sampling factor outside 1..4");
      }
      if (P.info.ncomp == 1) P.info.h[0] = P.info.v[0] = 1;      // a single-component image is never interleaved
      have_frame = true;
    } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
      return unsupported("lossless / hierarchical / arithmetic-coded JPEG");
    } else if (m == 0xC4) {
      size_t i = 0;
      while (i + 17 <= sl) {
        const int tc = s[i] >> 4, th = s[i] & 15;
        if (tc > 1 || th > 3) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad DHT id");
        HuffTable t;
        int nsym = 0;
        for (int k = 0; k < 16; ++k) { t.counts[k] = s[i + 1 + k]; nsym += t.counts[k]; }
        if (nsym > 256 || i + 17 + nsym > sl) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad DHT");
        memcpy(t.symbols, s + i + 17, nsym);
        if (!tc)
          for (int k = 0; k < nsym; ++k)
            if (t.symbols[k] > 15) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: DC Huffman symbol beyond 15");
        if (!build_huff(t)) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: DHT code lengths are not a prefix code");
        P.pool.push_back(t);
        (tc ? P.cur_ac : P.cur_dc)[th] = (int)P.pool.size() - 1;
        i += 17 + nsym;
      }
      if (i != sl) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad DHT length");
    } else if (m == 0xDD) {
      if (sl != 2) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad DRI length");
      ri = (s[0] << 8) | s[1];
    } else if (m == 0xE0 && sl >= 14 && !memcmp(s, "JFIF", 5)) {
      saw_jfif = true;
    } else if (m == 0xEE && sl >= 12 && !memcmp(s, "Adobe", 5)) {
      saw_adobe = true;
      adobe_transform = s[11];
    } else if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE || m == 0xDC) {
      // APPn / COM / DNL: skipped (jdmarker.c skip_variable)
    } else if (m != 0xDA) {
      return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: unexpected marker 0xFF%02X", (int)m);   // libjpeg: JERR_UNKNOWN_MARKER / SOI_DUPLICATE
    } else {
      if (!have_frame) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: SOS before SOF");
      Scan sc;
      memset(&sc, 0, sizeof(sc));
      sc.ns = sl >= 1 ? s[0] : 0;
      if (sc.ns < 1 || sc.ns > P.info.ncomp || sl != (size_t)(4 + 2 * sc.ns)) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad SOS");
      if (!P.info.progressive && sc.ns != P.info.ncomp) return unsupported("non-interleaved sequential (multi-scan) file");
      for (int c = 0; c < sc.ns; ++c) {
        int idx = -1;
        for (int k = 0; k < P.info.ncomp; ++k)
          if (comp_id[k] == s[1 + 2 * c]) idx = k;
        if (idx < 0) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: SOS names an unknown component");
        for (int k = 0; k < c; ++k)
          if (sc.ci[k] == idx) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: SOS names a component twice");
        sc.ci[c] = idx;
        sc.td[c] = s[2 + 2 * c] >> 4;
        sc.ta[c] = s[2 + 2 * c] & 15;
        if (sc.td[c] > 3 || sc.ta[c] > 3) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad table selector");
      }
      sc.ss = s[1 + 2 * sc.ns];
      sc.se = s[2 + 2 * sc.ns];
      sc.ah = s[3 + 2 * sc.ns] >> 4;
      sc.al = s[3 + 2 * sc.ns] & 15;
      if (!P.info.progressive) { sc.ss = 0; sc.se = 63; sc.ah = sc.al = 0; }
      if (sc.ss > sc.se || sc.se > 63 || sc.al > 13 || (sc.ss == 0 && sc.se != 0 && P.info.progressive) ||
          (sc.ss > 0 && sc.ns != 1) || (sc.ah != 0 && sc.al != sc.ah - 1))
        return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: bad progressive scan parameters");
      for (int c = 0; c < sc.ns; ++c)                 // jdphuff.c start_pass_phuff_decoder: precision each coefficient has reached
        for (int k = sc.ss; k <= sc.se; ++k) coef_bits[sc.ci[c]][k] = sc.al;
      for (int c = 0; c < sc.ns; ++c) {
        const bool need_dc = sc.ss == 0 && sc.ah == 0, need_ac = sc.se > 0;
        if ((need_dc && P.cur_dc[sc.td[c]] < 0) || (need_ac && P.cur_ac[sc.ta[c]] < 0))
          return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: scan refers to a missing Huffman table");
      }
      for (int k = 0; k < 4; ++k) { sc.dc_idx[k] = P.cur_dc[k]; sc.ac_idx[k] = P.cur_ac[k]; }
      sc.ri = ri;
      sc.ecs = pos + ln;
      P.scans.push_back(sc);
      if (!P.info.progressive) break;
    }
    pos += ln;
  }
  if (P.scans.empty()) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: no SOS marker");
  {                                              // Pillow refuses a file that ends before its EOI marker ("image file is truncated")
    bool eoi = false;
    for (size_t i = P.scans.back().ecs; i + 1 < n && !eoi; ++i) eoi = d[i] == 0xFF && d[i + 1] == 0xD9;
    if (!eoi) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: truncated file (no EOI marker after the last scan)");
  }
  // A progressive file that stops before every coefficient reached full precision: libjpeg would smooth the blocks
  // (jdcoefct.c smoothing_ok / decompress_smooth_data); this decoder leaves such files to the caller's host decoder.
  if (P.info.progressive)
    for (int c = 0; c < P.info.ncomp; ++c)
      for (int k = 0; k < 64; ++k)
        if (coef_bits[c][k] != 0) return unsupported("progressive file with an incomplete scan script");
  // colour space as libjpeg's default_decompress_parms decides it
  if (P.info.ncomp == 3) {
    bool ycc = true;
    if (saw_jfif) ycc = true;
    else if (saw_adobe) ycc = adobe_transform == 1;
    else if (comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B') ycc = false;
    if (!ycc) return unsupported("RGB-coded JPEG (no YCbCr transform)");
    if (P.info.h[1] != 1 || P.info.v[1] != 1 || P.info.h[2] != 1 || P.info.v[2] != 1)
      return unsupported("chroma sampling factors other than 1x1");
    if (!((P.info.h[0] == 1 && P.info.v[0] == 1) || (P.info.h[0] == 2 && P.info.v[0] == 1) || (P.info.h[0] == 2 && P.info.v[0] == 2)))
      return unsupported("luma sampling other than 1x1, 2x1, 2x2");
  }
  for (int c = 0; c < P.info.ncomp; ++c)
    if (!P.qt_present[P.info.tq[c]]) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_parse: missing quantisation table");
  const int hmax = P.info.h[0], vmax = P.info.v[0];
  const int mcux = (P.info.width + 8 * hmax - 1) / (8 * hmax), mcuy = (P.info.height + 8 * vmax - 1) / (8 * vmax);
  for (int c = 0; c < P.info.ncomp; ++c) {
    P.info.blocks_w[c] = mcux * P.info.h[c];
    P.info.blocks_h[c] = mcuy * P.info.v[c];
  }
  P.info.restart_interval = P.scans[0].ri;
  P.info.supported = 1;
  return 0;
}

struct BitReader {
  const uint8_t* d;
  size_t n, p;
  uint64_t acc;
  int cnt;
  bool hit_marker;
  int fake = 0;                                          // zero bits fed behind a marker / the end of the file (the tail of acc)
  int next_rst = 0;
  inline void fill() {
    if (!hit_marker && p + 8 <= n && cnt <= 56) {          // eight bytes at once when none of them is 0xFF
      uint64_t v;
      memcpy(&v, d + p, 8);
      if (((~v - 0x0101010101010101ull) & v & 0x8080808080808080ull) == 0) {
        v = __builtin_bswap64(v);
        const int take = (64 - cnt) >> 3;
        acc |= (v >> (64 - 8 * take)) << (64 - cnt - 8 * take);
        cnt += 8 * take;
        p += take;
        return;
      }
    }
    fill_bytewise();
  }
  void fill_bytewise() {
    while (cnt <= 56) {
      uint8_t b = 0;
      bool real = false;
      if (!hit_marker && p < n) {
        b = d[p++];
        real = true;
        if (b == 0xFF) {
          const uint8_t nx = p < n ? d[p] : 0xD9;
          if (nx == 0) ++p;
          else { hit_marker = true; --p; b = 0; real = false; }   // stay on the marker, feed zeros
        }
      }
      if (!real) fake += 8;
      acc |= (uint64_t)b << (56 - cnt);
      cnt += 8;
    }
  }
  // The segment ended exactly where the decoder stopped: no bit was taken from behind the marker and less than one byte of
  // (padding) bits is left in front of it.  libjpeg decodes such streams too, with a warning ("premature end of data segment",
  // "extraneous bytes before marker") and its own recovery; this decoder refuses them so that the caller's host decoder decides.
  bool clean_end() {
    if (cnt < fake || cnt - fake >= 8) return false;
    return hit_marker || (p + 1 < n && d[p] == 0xFF && d[p + 1] != 0);
  }
  inline uint32_t peek(int k) { return (uint32_t)(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; cnt -= k; }
  inline int get_bits(int k) {                           // k <= 16
    if (!k) return 0;
    if (cnt < k) fill();
    const int v = (int)peek(k);
    skip(k);
    return v;
  }
  inline int receive_extend(int t) {
    if (!t) return 0;
    if (cnt < t) fill();
    const int v = (int)peek(t);
    skip(t);
    return v < (1 << (t - 1)) ? v - ((1 << t) - 1) : v;
  }
  inline int decode(const HuffTable& t) {
    if (cnt < 16) fill();
    const uint16_t f = t.fast[peek(9)];
    if (f) { skip(f >> 8); return f & 255; }
    int code = (int)peek(9);
    for (int l = 10; l <= 16; ++l) {
      code = (int)peek(l);
      if (code <= t.maxcode[l] && t.maxcode[l] >= 0 && code >= t.mincode[l]) {
        skip(l);
        return t.symbols[t.valptr[l] + code - t.mincode[l]];
      }
    }
    return -1;
  }
  bool restart() {                                       // byte-align; the next marker must be the RSTn that is due
    if (!clean_end() || p + 1 >= n || d[p] != 0xFF || d[p + 1] != 0xD0 + (next_rst & 7)) return false;
    ++next_rst;
    acc = 0; cnt = 0; fake = 0; hit_marker = false;
    p += 2;
    return true;
  }
};

// ---- device side -------------------------------------------------------------------------------------------------
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jpeg_idct_islow: in[0..7] (stride s) -> out[0..7] (stride so), descaled by `shift`
__device__ __forceinline__ void idct_1d(const int* in, int s, int* out, int so, int shift) {
  int z2 = in[2 * s], z3 = in[6 * s];
  int z1 = (z2 + z3) * FIX_0_541196100;
  int tmp2 = z1 + z3 * (-FIX_1_847759065);
  int tmp3 = z1 + z2 * FIX_0_765366865;
  z2 = in[0];
  z3 = in[4 * s];
  int tmp0 = (z2 + z3) << 13, tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7 * s];
  tmp1 = in[5 * s];
  tmp2 = in[3 * s];
  tmp3 = in[1 * s];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * FIX_1_175875602;
  tmp0 *= FIX_0_298631336;
  tmp1 *= FIX_2_053119869;
  tmp2 *= FIX_3_072711026;
  tmp3 *= FIX_1_501321110;
  z1 *= -FIX_0_899976223;
  z2 *= -FIX_2_562915447;
  z3 = z3 * (-FIX_1_961570560) + z5;
  z4 = z4 * (-FIX_0_390180644) + z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  out[0] = descale(tmp10 + tmp3, shift);
  out[7 * so] = descale(tmp10 - tmp3, shift);
  out[1 * so] = descale(tmp11 + tmp2, shift);
  out[6 * so] = descale(tmp11 - tmp2, shift);
  out[2 * so] = descale(tmp12 + tmp1, shift);
  out[5 * so] = descale(tmp12 - tmp1, shift);
  out[3 * so] = descale(tmp13 + tmp0, shift);
  out[4 * so] = descale(tmp13 - tmp0, shift);
}

// one thread per 8x8 block: dequantise, columns pass (>> 11), rows pass (>> 18), +128, clamp -> plane [bh*8, bw*8]
__global__ void __launch_bounds__(128)
jpeg_idct_kernel(const int16_t* __restrict__ coefs, const uint16_t* __restrict__ qt, uint8_t* __restrict__ plane, int blocks_w,
                 int nblocks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  int d[64], ws[64];
  const int4* src = reinterpret_cast<const int4*>(coefs + (int64_t)b * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int4 v = src[i];
    const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      d[8 * i + 2 * k] = (int)(int16_t)(w[k] & 0xffff) * (int)qt[8 * i + 2 * k];
      d[8 * i + 2 * k + 1] = (int)(int16_t)(w[k] >> 16) * (int)qt[8 * i + 2 * k + 1];
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) idct_1d(d + c, 8, ws + c, 8, 13 - 2);
  const int by = b / blocks_w, bx = b - by * blocks_w;
  uint8_t* dst = plane + ((int64_t)by * 8) * ((int64_t)blocks_w * 8) + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int o[8];
    idct_1d(ws + 8 * r, 1, o, 1, 13 + 2 + 3);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= (uint32_t)min(max(o[k] + 128, 0), 255) << (8 * k);
      hi |= (uint32_t)min(max(o[4 + k] + 128, 0), 255) << (8 * k);
    }
    *reinterpret_cast<uint2*>(dst + (int64_t)r * blocks_w * 8) = make_uint2(lo, hi);
  }
}

struct ColorArgs {
  const uint8_t* y;
  const uint8_t* cb;
  const uint8_t* cr;
  int ys, cs;              // row strides of the luma / chroma planes
  int W, H, wd, hd;        // image size, valid chroma size (ceil(W / 2), ceil(H / 2) where subsampled)
  int mode;                // 0 = gray, 1 = 4:4:4, 2 = h2v1 fancy, 3 = h2v2 fancy, 4 = h2v1 replicate, 5 = h2v2 replicate
};

__device__ __forceinline__ int chroma_at(const uint8_t* c, int cs, int x, int y, const ColorArgs& a) {
  if (a.mode == 1) return c[(int64_t)y * cs + x];
  const int cx = x >> 1;
  if (a.mode == 4) return c[(int64_t)y * cs + cx];            // jdsample.c h2v1_upsample / h2v2_upsample: plain replication
  if (a.mode == 5) return c[(int64_t)(y >> 1) * cs + cx];
  if (a.mode == 2) {                                          // jdsample.c h2v1_fancy_upsample
    const int p = c[(int64_t)y * cs + cx];
    if (x & 1) return x == 2 * a.wd - 1 ? p : (3 * p + c[(int64_t)y * cs + cx + 1] + 2) >> 2;
    return x == 0 ? p : (3 * p + c[(int64_t)y * cs + cx - 1] + 1) >> 2;
  }
  // h2v2_fancy_upsample: 3:1 blend with the nearer neighbouring chroma row (edge rows replicated), then the triangle in x
  const int cy = y >> 1;
  int oy = (y & 1) ? cy + 1 : cy - 1;
  oy = oy < 0 ? 0 : (oy > a.hd - 1 ? a.hd - 1 : oy);
  const uint8_t* r0 = c + (int64_t)cy * cs;
  const uint8_t* r1 = c + (int64_t)oy * cs;
  const int s = 3 * r0[cx] + r1[cx];
  if (x & 1) return x == 2 * a.wd - 1 ? (4 * s + 7) >> 4 : (3 * s + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  return x == 0 ? (4 * s + 8) >> 4 : (3 * s + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// one thread per output pixel: upsample + jdcolor.c ycc_rgb_convert -> interleaved RGB bytes [H, W, 3]
__global__ void __launch_bounds__(256)
jpeg_color_kernel(const ColorArgs a, uint8_t* __restrict__ rgb) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)a.W * a.H) return;
  const int y = (int)(idx / a.W), x = (int)(idx - (int64_t)y * a.W);
  const int Y = a.y[(int64_t)y * a.ys + x];
  int r = Y, g = Y, b = Y;
  if (a.mode != 0) {
    const int cb = chroma_at(a.cb, a.cs, x, y, a) - 128, cr = chroma_at(a.cr, a.cs, x, y, a) - 128;
    r = Y + ((91881 * cr + 32768) >> 16);                        // FIX(1.40200)
    b = Y + ((116130 * cb + 32768) >> 16);                       // FIX(1.77200)
    g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);          // FIX(0.34414), FIX(0.71414)
    r = min(max(r, 0), 255);
    g = min(max(g, 0), 255);
    b = min(max(b, 0), 255);
  }
  uint8_t* o = rgb + idx * 3;
  o[0] = (uint8_t)r;
  o[1] = (uint8_t)g;
  o[2] = (uint8_t)b;
}

size_t plane_bytes(const advgrpo_jpeg_info& i, int c) { return ((size_t)i.blocks_w[c] * 8 * i.blocks_h[c] * 8 + 255) & ~(size_t)255; }

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

int advgrpo_jpeg_parse(const uint8_t* file, size_t nbytes, advgrpo_jpeg_info* info) {
  ADVGRPO_CHECK_ARG(file && info, "jpeg_parse: null pointer");
  Parsed* P = new Parsed();
  const int rc = parse_jpeg(file, nbytes, *P);
  *info = P->info;
  delete P;
  if (rc == 1) { info->supported = 0; return ADVGRPO_OK; }      // a valid file this decoder does not take: not an error
  return rc;
}

size_t advgrpo_jpeg_coef_count(const advgrpo_jpeg_info* info) {
  if (!info || !info->supported) return 0;
  size_t n = 0;
  for (int c = 0; c < info->ncomp; ++c) n += (size_t)info->blocks_w[c] * info->blocks_h[c] * 64;
  return n;
}

namespace {

// sequential scan: one interleaved pass over the MCUs
int decode_sequential(const uint8_t* file, size_t nbytes, const Parsed& P, int16_t* coefs, const size_t* off) {
  const advgrpo_jpeg_info& I = P.info;
  const Scan& sc = P.scans[0];
  BitReader br{file, nbytes, sc.ecs, 0, 0, false};
  int pred[3] = {0, 0, 0};
  const int mcux = I.blocks_w[0] / I.h[0], mcuy = I.blocks_h[0] / I.v[0];
  int n = 0;
  for (int my = 0; my < mcuy; ++my)
    for (int mx = 0; mx < mcux; ++mx) {
      if (sc.ri && n && n % sc.ri == 0) {
        if (!br.restart()) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: restart marker not where it is due");
        pred[0] = pred[1] = pred[2] = 0;
      }
      ++n;
      for (int k = 0; k < sc.ns; ++k) {
        const int c = sc.ci[k];
        const HuffTable& dct = P.pool[sc.dc_idx[sc.td[k]]];
        const HuffTable& act = P.pool[sc.ac_idx[sc.ta[k]]];
        const uint16_t* qt = P.qt[I.tq[c]];
        for (int by = 0; by < I.v[c]; ++by)
          for (int bx = 0; bx < I.h[c]; ++bx) {
            int16_t* blk = coefs + off[c] + ((size_t)(my * I.v[c] + by) * I.blocks_w[c] + (mx * I.h[c] + bx)) * 64;
            memset(blk, 0, 64 * sizeof(int16_t));
            int col[8] = {0, 0, 0, 0, 0, 0, 0, 0};                 // column sums of |coefficient x quantiser| (kMaxColumnSum)
            auto put = [&](int nat, int v) {
              blk[nat] = (int16_t)v;
              const int m = (int)(int16_t)v * (int)qt[nat];
              col[nat & 7] += m < 0 ? -m : m;
            };
            const int t = br.decode(dct);
            if (t < 0 || t > 11) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: corrupt DC code");
            pred[c] += br.receive_extend(t);
            put(0, pred[c]);
            for (int kk = 1; kk < 64;) {
              if (br.cnt < 16) br.fill();
              const int32_t fa = act.fast_ac[br.peek(10)];
              if (fa) {                                           // code and magnitude bits in one lookup
                kk += (fa >> 8) & 15;
                if (kk > 63) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: coefficient index out of range");
                put(kZigzag[kk++], fa >> 16);
                br.skip(fa & 255);
                continue;
              }
              const int rs = br.decode(act);
              if (rs < 0) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: corrupt AC code");
              const int r = rs >> 4, sz = rs & 15;
              if (sz == 0) {
                if (r != 15) break;
                kk += 16;
                continue;
              }
              kk += r;
              if (kk > 63) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: coefficient index out of range");
              put(kZigzag[kk], br.receive_extend(sz));
              ++kk;
            }
            for (int q = 0; q < 8; ++q)
              if (col[q] > kMaxColumnSum)
                return set_error(ADVGRPO_ERR_UNSUPPORTED, "jpeg_entropy_decode: coefficients beyond the range of 8-bit samples");
          }
      }
    }
  if (!br.clean_end()) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: the entropy-coded segment does not end with the last MCU");
  size_t q = br.p;                                       // on the marker that ended the segment: it must be EOI (fill bytes allowed)
  while (q + 1 < nbytes && file[q] == 0xFF && file[q + 1] == 0xFF) ++q;
  if (q + 1 >= nbytes || file[q] != 0xFF || file[q + 1] != 0xD9)
    return set_error(ADVGRPO_ERR_UNSUPPORTED, "jpeg_entropy_decode: more segments behind the scan of a sequential file");
  return ADVGRPO_OK;
}

// progressive file (ITU T.81 annex G, libjpeg jdphuff.c): the scans accumulate into the coefficient blocks
int decode_progressive(const uint8_t* file, size_t nbytes, const Parsed& P, int16_t* coefs, const size_t* off) {
  const advgrpo_jpeg_info& I = P.info;
  const int hmax = I.h[0], vmax = I.v[0];
  const int mcux = I.blocks_w[0] / hmax, mcuy = I.blocks_h[0] / vmax;
  // Per block, the zigzag index of the last AC coefficient written so far: everything behind it is still zero, so the
  // refinement scans (which walk a block's band once per scan) stop there instead of reading 63 coefficients of every block.
  size_t total_blocks = 0;
  for (int c = 0; c < I.ncomp; ++c) total_blocks += (size_t)I.blocks_w[c] * I.blocks_h[c];
  std::vector<uint8_t> last_nz(total_blocks, 0);
  for (const Scan& sc : P.scans) {
    BitReader br{file, nbytes, sc.ecs, 0, 0, false};
    int pred[3] = {0, 0, 0};
    int eobrun = 0;
    // units: MCUs of an interleaved scan, or the component's own blocks (ceil(w_c / 8) x ceil(h_c / 8)) otherwise
    int ux_n = mcux, uy_n = mcuy;
    if (sc.ns == 1) {
      const int c = sc.ci[0];
      const int wc = (I.width * I.h[c] + hmax - 1) / hmax, hc = (I.height * I.v[c] + vmax - 1) / vmax;
      ux_n = (wc + 7) / 8;
      uy_n = (hc + 7) / 8;
    }
    const int p1 = 1 << sc.al, m1 = -(1 << sc.al);
    int n = 0;
    for (int uy = 0; uy < uy_n; ++uy)
      for (int ux = 0; ux < ux_n; ++ux) {
        if (sc.ri && n && n % sc.ri == 0) {
          if (!br.restart()) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: restart marker not where it is due");
          pred[0] = pred[1] = pred[2] = 0;
          eobrun = 0;
        }
        ++n;
        for (int k = 0; k < sc.ns; ++k) {
          const int c = sc.ci[k];
          const int nby = sc.ns == 1 ? 1 : I.v[c], nbx = sc.ns == 1 ? 1 : I.h[c];
          for (int by = 0; by < nby; ++by)
            for (int bx = 0; bx < nbx; ++bx) {
              const int brow = sc.ns == 1 ? uy : uy * I.v[c] + by, bcol = sc.ns == 1 ? ux : ux * I.h[c] + bx;
              int16_t* blk = coefs + off[c] + ((size_t)brow * I.blocks_w[c] + bcol) * 64;
              uint8_t& last = last_nz[(size_t)(blk - coefs) / 64];
              if (sc.ss == 0) {                                   // DC scan
                if (sc.ah == 0) {
                  const int t = br.decode(P.pool[sc.dc_idx[sc.td[k]]]);
                  if (t < 0 || t > 11) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: corrupt DC code");
                  pred[c] += br.receive_extend(t);
                  blk[0] = (int16_t)(pred[c] * (1 << sc.al));
                } else if (br.get_bits(1)) {
                  blk[0] = (int16_t)(blk[0] | p1);
                }
                continue;
              }
              const HuffTable& act = P.pool[sc.ac_idx[sc.ta[k]]];
              if (sc.ah == 0) {                                   // AC first scan
                if (eobrun > 0) { --eobrun; continue; }
                for (int kk = sc.ss; kk <= sc.se;) {
                  if (br.cnt < 16) br.fill();
                  const int32_t fa = act.fast_ac[br.peek(10)];
                  if (fa) {
                    kk += (fa >> 8) & 15;
                    if (kk > 63) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: coefficient index out of range");
                    if (kk > last) last = (uint8_t)kk;
                    blk[kZigzag[kk++]] = (int16_t)((fa >> 16) * (1 << sc.al));
                    br.skip(fa & 255);
                    continue;
                  }
                  const int rs = br.decode(act);
                  if (rs < 0) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: corrupt AC code");
                  const int r = rs >> 4, sz = rs & 15;
                  if (sz) {
                    kk += r;
                    if (kk > 63) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: coefficient index out of range");
                    if (kk > last) last = (uint8_t)kk;
                    blk[kZigzag[kk]] = (int16_t)(br.receive_extend(sz) * (1 << sc.al));
                    ++kk;
                  } else if (r == 15) {
                    kk += 16;
                  } else {
                    eobrun = (1 << r) + br.get_bits(r) - 1;
                    break;
                  }
                }
                continue;
              }
              // AC refinement scan (jdphuff.c decode_mcu_AC_refine)
              int kk = sc.ss;
              if (eobrun == 0) {
                while (kk <= sc.se) {
                  const int rs = br.decode(act);
                  if (rs < 0) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: corrupt AC code");
                  int r = rs >> 4;
                  const int sz = rs & 15;
                  int val = 0;
                  if (sz) {
                    val = br.get_bits(1) ? p1 : m1;                // the size is always 1 in a refinement scan
                  } else if (r != 15) {
                    eobrun = (1 << r) + br.get_bits(r);
                    break;
                  }
                  while (kk <= sc.se) {                            // skip r still-zero coefficients, refining the others
                    if (kk > last) {                               // nothing but zeros from here on: jump
                      kk = r > sc.se - kk ? sc.se + 1 : kk + r;
                      break;
                    }
                    int16_t& co = blk[kZigzag[kk]];
                    if (co != 0) {
                      if (br.get_bits(1) && (co & p1) == 0) co = (int16_t)(co + (co >= 0 ? p1 : m1));
                    } else {
                      if (r == 0) break;
                      --r;
                    }
                    ++kk;
                  }
                  if (val) {                                       // a new coefficient behind the band: only a corrupt stream does that
                    if (kk > sc.se) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: refinement scan runs out of its band");
                    if (kk > last) last = (uint8_t)kk;
                    blk[kZigzag[kk]] = (int16_t)val;
                  }
                  ++kk;
                }
              }
              if (eobrun > 0) {                                    // the rest of the band: correction bits only
                for (const int end = sc.se < last ? sc.se : last; kk <= end; ++kk) {
                  int16_t& co = blk[kZigzag[kk]];
                  if (co != 0 && br.get_bits(1) && (co & p1) == 0) co = (int16_t)(co + (co >= 0 ? p1 : m1));
                }
                --eobrun;
              }
            }
        }
      }
    if (!br.clean_end()) return set_error(ADVGRPO_ERR_BAD_ARG, "jpeg_entropy_decode: a scan's entropy-coded segment does not end with its last block");
  }
  return ADVGRPO_OK;
}

}  // namespace

int advgrpo_jpeg_entropy_decode(const uint8_t* file, size_t nbytes, int16_t* coefs_host, uint16_t* qtabs_host) {
  ADVGRPO_CHECK_ARG(file && coefs_host && qtabs_host, "jpeg_entropy_decode: null pointer");
  Parsed* P = new Parsed();
  int rc = parse_jpeg(file, nbytes, *P);
  if (rc != 0) {
    delete P;
    return rc == 1 ? set_error(ADVGRPO_ERR_UNSUPPORTED, "jpeg_entropy_decode: file is outside the supported subset") : rc;
  }
  const advgrpo_jpeg_info& I = P->info;
  size_t off[3] = {0, 0, 0}, total = 0;
  for (int c = 0; c < I.ncomp; ++c) {
    off[c] = total;
    total += (size_t)I.blocks_w[c] * I.blocks_h[c] * 64;
    for (int k = 0; k < 64; ++k) qtabs_host[c * 64 + k] = P->qt[I.tq[c]][k];
  }
  if (!I.progressive) {
    rc = decode_sequential(file, nbytes, *P, coefs_host, off);       // zeroes each block and checks its range as it goes
  } else {
    memset(coefs_host, 0, total * sizeof(int16_t));
    rc = decode_progressive(file, nbytes, *P, coefs_host, off);
    for (int c = 0; c < I.ncomp && rc == ADVGRPO_OK; ++c) {             // the scans accumulate: range check (kMaxColumnSum) at the end
      const uint16_t* q = P->qt[I.tq[c]];
      const int16_t* blk = coefs_host + off[c];
      const size_t nb = (size_t)I.blocks_w[c] * I.blocks_h[c];
      for (size_t b = 0; b < nb && rc == ADVGRPO_OK; ++b, blk += 64) {
        int col[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < 8; ++r)
          for (int k = 0; k < 8; ++k) {
            const int v = (int)blk[8 * r + k] * (int)q[8 * r + k];
            col[k] += v < 0 ? -v : v;
          }
        for (int k = 0; k < 8; ++k)
          if (col[k] > kMaxColumnSum)
            rc = set_error(ADVGRPO_ERR_UNSUPPORTED, "jpeg_entropy_decode: coefficients beyond the range of 8-bit samples");
      }
    }
  }
  delete P;
  return rc;
}

size_t advgrpo_jpeg_workspace_bytes(const advgrpo_jpeg_info* info) {
  if (!info || !info->supported) return 0;
  size_t n = 256;
  for (int c = 0; c < info->ncomp; ++c) n += plane_bytes(*info, c);
  return n;
}

int advgrpo_jpeg_idct_to_rgb(const int16_t* coefs_dev, const uint16_t* qtabs_dev, const advgrpo_jpeg_info* info,
                             uint8_t* rgb_hwc_dev, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(coefs_dev && qtabs_dev && info && rgb_hwc_dev, "jpeg_idct_to_rgb: null pointer");
  ADVGRPO_CHECK_ARG(info->supported && (info->ncomp == 1 || info->ncomp == 3) && info->width >= 1 && info->height >= 1,
                    "jpeg_idct_to_rgb: unsupported file (advgrpo_jpeg_parse reported supported = 0)");
  ADVGRPO_CHECK_ARG(aligned16(coefs_dev), "jpeg_idct_to_rgb: coefficients must be 16-byte aligned");
  if (!workspace || workspace_bytes < advgrpo_jpeg_workspace_bytes(info))
    return set_error(ADVGRPO_ERR_WORKSPACE, "jpeg_idct_to_rgb: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* planes[3] = {nullptr, nullptr, nullptr};
  size_t woff = 0, coff = 0;
  for (int c = 0; c < info->ncomp; ++c) {
    planes[c] = (uint8_t*)workspace + woff;
    woff += plane_bytes(*info, c);
    const int nb = info->blocks_w[c] * info->blocks_h[c];
    jpeg_idct_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(coefs_dev + coff, qtabs_dev + c * 64, planes[c],
                                                                  info->blocks_w[c], nb);
    ADVGRPO_CUDA_LAUNCH_CHECK();
    coff += (size_t)nb * 64;
  }
  ColorArgs a;
  a.y = planes[0];
  a.cb = planes[1];
  a.cr = planes[2];
  a.ys = info->blocks_w[0] * 8;
  a.cs = info->ncomp == 3 ? info->blocks_w[1] * 8 : 0;
  a.W = info->width;
  a.H = info->height;
  a.mode = info->ncomp == 1 ? 0 : (info->h[0] == 1 ? 1 : (info->v[0] == 1 ? 2 : 3));
  a.wd = a.mode >= 2 ? (info->width + 1) / 2 : info->width;
  a.hd = a.mode == 3 ? (info->height + 1) / 2 : info->height;
  if (a.mode >= 2 && a.wd <= 2) a.mode += 2;   // jinit_upsampler: the fancy (triangle) filters only when downsampled_width > 2
  const int64_t npx = (int64_t)a.W * a.H;
  jpeg_color_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(a, rgb_hwc_dev);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"
