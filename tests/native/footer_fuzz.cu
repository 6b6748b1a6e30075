// footer_fuzz.cu -- CPU-only robustness check of the engine's Parquet footer reader: mutated and truncated footers must
// be rejected with an hs::Error (or parse to something), never crash or hang.  argv[1] = a valid Parquet file,
// argv[2] = number of mutations.  Deterministic (splitmix64 stream).
#include <cstdio>
#include <cstring>

#include "../../hyperspace_b200/csrc/parquet_meta.h"

using namespace hs;

static uint64_t next(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* in = fopen(argv[1], "rb");
  if (!in) return 3;
  std::vector<uint8_t> file;
  uint8_t buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, in)) > 0) file.insert(file.end(), buf, buf + n);
  fclose(in);
  uint32_t flen;
  memcpy(&flen, file.data() + file.size() - 8, 4);
  const size_t foot0 = file.size() - 8 - flen;
  const int rounds = atoi(argv[2]);
  uint64_t seed = 7;
  int rejected = 0, accepted = 0;
  for (int r = 0; r < rounds; r++) {
    std::vector<uint8_t> m = file;
    const int kind = (int)(next(seed) % 4);
    if (kind == 0) {  // flip 1..4 bytes inside the footer
      const int k = 1 + (int)(next(seed) % 4);
      for (int i = 0; i < k; i++) m[foot0 + next(seed) % flen] = (uint8_t)next(seed);
    } else if (kind == 1) {  // lie about the footer length
      uint32_t bad = (uint32_t)next(seed);
      if (next(seed) % 2) bad %= (uint32_t)(2 * m.size());
      memcpy(m.data() + m.size() - 8, &bad, 4);
    } else if (kind == 2) {  // truncate the footer and re-seal it
      const uint32_t keep = (uint32_t)(next(seed) % flen);
      m.resize(foot0 + keep);
      m.insert(m.end(), (const uint8_t*)&keep, (const uint8_t*)&keep + 4);
      m.insert(m.end(), {'P', 'A', 'R', '1'});
    } else {  // overwrite a run with 0xff / 0x00 (long varints, huge list sizes)
      const size_t at = foot0 + next(seed) % flen, len = 1 + next(seed) % 12;
      const uint8_t v = next(seed) % 2 ? 0xff : 0x00;
      for (size_t i = at; i < std::min(at + len, foot0 + flen); i++) m[i] = v;
    }
    if (getenv("FUZZ_TRACE")) {
      fprintf(stderr, "round %d kind %d\n", r, kind);
      fflush(stderr);
    }
    try {
      const pq::FileMeta fm = pq::parse_footer(m.data(), m.size(), "fuzz");
      (void)fm;
      accepted++;
    } catch (const Error&) {
      rejected++;
    } catch (const std::bad_alloc&) {
      printf("bad_alloc at round %d (a length field was trusted)\n", r);
      return 1;
    } catch (const std::length_error&) {
      printf("length_error at round %d (a length field was trusted)\n", r);
      return 1;
    }
  }
  printf("rounds=%d rejected=%d accepted=%d\n", rounds, rejected, accepted);
  return 0;
}
