// AddressSanitizer harness for the host halves of the image decoders (advgrpo_png_parse / _inflate, advgrpo_jpeg_parse /
// _entropy_decode): every seed file given on the command line is mutated N times (truncation, byte / bit flips, deletions,
// splices) and pushed through the C-ABI with output buffers of EXACTLY the size the library asks for, so any write or read
// outside them aborts.  Built and run by tests/test_decoder_fuzz.py (CPU only; no CUDA call is made).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "advgrpo_b200.h"

static uint64_t s = 88172645463325252ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

static std::vector<uint8_t> mutate(const std::vector<uint8_t>& in) {
  std::vector<uint8_t> b = in;
  const int kinds = 1 + (int)(rnd() % 3);
  for (int q = 0; q < kinds && b.size() > 2; ++q) {
    switch (rnd() % 5) {
      case 0: b.resize(1 + rnd() % b.size()); break;
      case 1: for (int k = 0, n = 1 + (int)(rnd() % 6); k < n; ++k) b[rnd() % b.size()] = (uint8_t)rnd(); break;
      case 2: b[rnd() % b.size()] ^= (uint8_t)(1u << (rnd() % 8)); break;
      case 3: { const size_t i = rnd() % b.size(), j = i + rnd() % 40; b.erase(b.begin() + i, b.begin() + (j < b.size() ? j : b.size())); break; }
      default: { const size_t i = rnd() % b.size(), j = rnd() % b.size(), n = rnd() % 32;          // splice a run from elsewhere
                 for (size_t k = 0; k < n && i + k < b.size() && j + k < b.size(); ++k) b[i + k] = b[j + k]; }
    }
  }
  return b;
}

// The PNG decoder verifies every chunk checksum, so most mutants would stop there: repair the checksums of half of them to let
// the inflate and scan-line checks see corrupt data too.
static uint32_t crc32(const uint8_t* p, size_t n) {
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) {
    c ^= p[i];
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
  }
  return ~c;
}

static void repair_png_checksums(std::vector<uint8_t>& b) {
  size_t pos = 8;
  while (pos + 12 <= b.size()) {
    const size_t ln = ((size_t)b[pos] << 24) | (b[pos + 1] << 16) | (b[pos + 2] << 8) | b[pos + 3];
    if (ln > b.size() || pos + 12 + ln > b.size()) break;
    const uint32_t c = crc32(&b[pos + 4], 4 + ln);
    for (int k = 0; k < 4; ++k) b[pos + 8 + ln + k] = (uint8_t)(c >> (24 - 8 * k));
    pos += 12 + ln;
  }
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s N seed-file...\n", argv[0]); return 2; }
  const int N = atoi(argv[1]);
  long ok = 0, bad = 0;
  for (int f = 2; f < argc; ++f) {
    FILE* fp = fopen(argv[f], "rb");
    if (!fp) { perror(argv[f]); return 2; }
    std::vector<uint8_t> seed;
    uint8_t buf[4096];
    for (size_t n; (n = fread(buf, 1, sizeof(buf), fp)) > 0;) seed.insert(seed.end(), buf, buf + n);
    fclose(fp);
    const bool is_png = seed.size() > 4 && seed[1] == 'P';
    for (int it = 0; it <= N; ++it) {
      std::vector<uint8_t> m = it == 0 ? seed : mutate(seed);
      if (is_png && (it & 1)) repair_png_checksums(m);
      uint8_t* file = (uint8_t*)malloc(m.size() ? m.size() : 1);          // exact-size copy: reads past the end are caught too
      memcpy(file, m.data(), m.size());
      int rc;
      if (is_png) {
        advgrpo_png_info info;
        rc = advgrpo_png_parse(file, m.size(), &info);
        if (rc == 0 && info.supported) {
          const size_t raw = advgrpo_png_raw_bytes(&info);
          uint8_t* out = (uint8_t*)malloc(raw ? raw : 1);
          uint8_t pal[768];
          rc = advgrpo_png_inflate(file, m.size(), out, pal);
          free(out);
        }
      } else {
        advgrpo_jpeg_info info;
        rc = advgrpo_jpeg_parse(file, m.size(), &info);
        if (rc == 0 && info.supported) {
          const size_t n = advgrpo_jpeg_coef_count(&info);
          int16_t* co = (int16_t*)malloc(n ? n * 2 : 2);
          uint16_t qt[192];
          rc = advgrpo_jpeg_entropy_decode(file, m.size(), co, qt);
          free(co);
        }
      }
      if (it == 0 && rc != 0) { fprintf(stderr, "seed %s does not decode: %s\n", argv[f], advgrpo_last_error()); return 1; }
      (rc == 0 ? ok : bad)++;
      free(file);
    }
  }
  printf("decoded %ld rejected %ld\n", ok, bad);
  return 0;
}
