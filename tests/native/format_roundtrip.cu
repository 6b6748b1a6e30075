// format_roundtrip.cu -- CPU-only check of the engine's Parquet format layer (thrift_compact.h + parquet_meta.h).
//
// Builds one small index-shaped Parquet file entirely on the HOST with the same writers the engine uses for page
// headers, definition levels, dictionary pages and the footer (the GPU normally fills in the page bodies; here plain host
// loops do), writes it to argv[1], parses its footer back with the engine's own reader and prints what it found.
// tests/test_native_format.py then lets pyarrow -- an independent Parquet implementation -- read the file.
// No CUDA call is made: the program runs without a GPU.
#include <cstdio>
#include <cstring>

#include "../../hyperspace_b200/csrc/parquet_meta.h"

using namespace hs;

static void put(std::vector<uint8_t>& f, const void* p, size_t n) { f.insert(f.end(), (const uint8_t*)p, (const uint8_t*)p + n); }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const int64_t N = 1000, RG = 400, P = 150;  // rows, rows per row group, rows per page (ragged on purpose)
  std::vector<pq::SchemaColumn> schema(3);
  const char* names[3] = {"k", "d", "x"};
  const int32_t types[3] = {pq::INT64, pq::INT32, pq::DOUBLE};
  const int widths[3] = {8, 4, 8};
  for (int c = 0; c < 3; c++) {
    schema[c].name = names[c];
    schema[c].type = types[c];
    schema[c].repetition = pq::OPTIONAL;
  }
  auto k_of = [](int64_t i) { return (int64_t)(i * 7 - 300); };              // sorted, like an index key
  const int32_t dict[5] = {10, 20, 30, 40, 50};
  auto d_of = [](int64_t i) { return (uint32_t)((i * 3) % 5); };             // dictionary code of row i
  auto x_null = [](int64_t i) { return i % 7 == 3; };                        // nullable column
  auto x_of = [](int64_t i) { return (double)i * 0.25; };

  std::vector<uint8_t> f;
  put(f, "PAR1", 4);
  std::vector<pq::OutRowGroup> rgs;
  for (int64_t r0 = 0; r0 < N; r0 += RG) {
    const int64_t r1 = std::min(N, r0 + RG);
    pq::OutRowGroup g;
    g.num_rows = r1 - r0;
    g.file_offset = (int64_t)f.size();
    const size_t rg_begin = f.size();
    for (int c = 0; c < 3; c++) {
      pq::OutChunk ch;
      ch.type = types[c];
      ch.num_values = r1 - r0;
      ch.value_width = widths[c];
      ch.data_page_offset = (int64_t)f.size();
      const size_t chunk_begin = f.size();
      if (c == 1) {  // dictionary page first
        ch.has_dictionary = true;
        ch.dictionary_page_offset = (int64_t)f.size();
        pq::write_dict_page_header(f, (int32_t)sizeof dict, 5);
        put(f, dict, sizeof dict);
        ch.data_page_offset = (int64_t)f.size();
      }
      int64_t nulls = 0;
      for (int64_t p0 = r0; p0 < r1; p0 += P) {
        const int64_t np = std::min(P, r1 - p0);
        if (c == 0) {  // PLAIN, all valid; the prefix picks a definition-level split that 8-byte-aligns the values
          pq::write_plain_page_prefix(f, f.size(), np, 8);
          for (int64_t i = p0; i < p0 + np; i++) {
            const int64_t v = k_of(i);
            put(f, &v, 8);
          }
        } else if (c == 1) {  // PLAIN_DICTIONARY: one bit-packed run of 3-bit codes, LSB first
          const uint32_t bw = 3;
          pq::write_dict_data_page_prefix(f, np, bw);
          const int64_t groups = (np + 7) / 8;
          std::vector<uint8_t> body((size_t)groups * bw, 0);
          for (int64_t j = 0; j < np; j++) {
            const uint64_t bit = (uint64_t)j * bw;
            const uint32_t code = d_of(p0 + j);
            for (uint32_t b = 0; b < bw; b++)
              if (code & (1u << b)) body[(bit + b) >> 3] |= (uint8_t)(1u << ((bit + b) & 7));
          }
          put(f, body.data(), body.size());
        } else {  // nullable PLAIN: bit-packed definition levels + dense values
          int64_t nn = 0;
          for (int64_t i = p0; i < p0 + np; i++) nn += x_null(i) ? 0 : 1;
          nulls += np - nn;
          pq::write_nullable_page_prefix(f, np, nn, 8);
          std::vector<uint8_t> levels((size_t)((np + 7) / 8), 0);
          for (int64_t j = 0; j < np; j++)
            if (!x_null(p0 + j)) levels[j >> 3] |= (uint8_t)(1u << (j & 7));
          put(f, levels.data(), levels.size());
          for (int64_t i = p0; i < p0 + np; i++)
            if (!x_null(i)) {
              const double v = x_of(i);
              put(f, &v, 8);
            }
        }
      }
      ch.total_size = (int64_t)(f.size() - chunk_begin);
      ch.null_count = c == 2 ? nulls : 0;
      if (c == 0) {  // min / max of the sorted key: first and last value of the row group (patched on the GPU normally)
        ch.has_minmax = true;
        const int64_t lo = k_of(r0), hi = k_of(r1 - 1);
        memcpy(ch.min_le, &lo, 8);
        memcpy(ch.max_le, &hi, 8);
      }
      g.chunks.push_back(ch);
    }
    g.total_byte_size = (int64_t)(f.size() - rg_begin);
    rgs.push_back(g);
  }
  std::vector<pq::StatSlot> slots;
  std::vector<uint8_t> footer = pq::write_footer(schema, rgs, N, pq::spark_schema_json(schema), &slots);
  put(f, footer.data(), footer.size());
  const uint32_t flen = (uint32_t)footer.size();
  put(f, &flen, 4);
  put(f, "PAR1", 4);
  FILE* out = fopen(argv[1], "wb");
  if (!out || fwrite(f.data(), 1, f.size(), out) != f.size()) return 3;
  fclose(out);

  // Page bodies of real index files (pages of >= 4096 rows) must start 8-byte aligned wherever the page begins: the GPU
  // writes them with full-width stores.  (Tiny pages like the ones above may not have an aligned split; the engine then
  // uses its unaligned store path.)
  for (int64_t n : {(int64_t)4096, (int64_t)100000, (int64_t)131072})
    for (int W : {4, 8})
      for (uint64_t off = 0; off < 8; off++) {
        std::vector<uint8_t> pre;
        pq::write_plain_page_prefix(pre, 1000 + off, n, W);
        if ((1000 + off + pre.size()) % 8 != 0) {
          fprintf(stderr, "page body not aligned: n=%lld W=%d offset mod 8 = %llu\n", (long long)n, W, (unsigned long long)off);
          return 1;
        }
      }
  printf("aligned_prefixes=ok\n");

  // read it back with the engine's own footer parser
  const pq::FileMeta fm = pq::parse_footer(f.data(), f.size(), argv[1]);
  printf("rows=%lld row_groups=%zu columns=%zu stat_slots=%zu\n", (long long)fm.num_rows, fm.row_groups.size(), fm.columns.size(),
         slots.size());
  for (size_t c = 0; c < fm.columns.size(); c++)
    printf("column %s type=%d optional=%d\n", fm.columns[c].name.c_str(), fm.columns[c].type, fm.columns[c].repetition == pq::OPTIONAL);
  for (size_t g = 0; g < fm.row_groups.size(); g++)
    for (size_t c = 0; c < fm.row_groups[g].columns.size(); c++) {
      const pq::ColumnChunkMeta& cm = fm.row_groups[g].columns[c];
      printf("rg %zu col %zu values=%lld start=%lld bytes=%lld codec=%d\n", g, c, (long long)cm.num_values, (long long)cm.start(),
             (long long)cm.total_compressed_size, cm.codec);
    }
  // the statistics placeholders the GPU patches: their recorded offsets must point at the bytes written above
  for (const pq::StatSlot& s : slots) {
    int64_t lo, hi;
    memcpy(&lo, footer.data() + s.min_off[1], 8);
    memcpy(&hi, footer.data() + s.max_off[0], 8);
    printf("stat rg %d col %d min=%lld max=%lld\n", s.row_group, s.column, (long long)lo, (long long)hi);
  }
  return 0;
}
