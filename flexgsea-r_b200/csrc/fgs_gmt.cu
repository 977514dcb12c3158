// Gene-set ingest, native and done once: GMT text -> CSR of 1-based gene indices.
//
// Replaces, for the data format that feeds the hot path:
//   read_gmt            R/flexgsea.R:966-973  name <tab> description <tab> gene <tab> gene ...
//   filter_gene_sets    R/flexgsea.R:939-958  drop genes absent from the data, then sets whose
//                                             remaining size is outside [gs.size.min, gs.size.max]
//   match(set, gene.names)  :279, :412        first occurrence wins for duplicated gene names,
//                                             duplicates inside a set are kept (m counts them)
// Host code only (no device work); the CSR it returns is what fgs_set_gene_sets takes.
#include "flexgsea_b200.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace fgs { void set_error(const char *fmt, ...); }

struct fgs_gmt {
  std::vector<int> offsets;        // [n_sets + 1]
  std::vector<int> index;          // [nnz], 1-based
  std::vector<std::string> names;  // kept sets, file order
  long long genes_before = 0, genes_after = 0;
  int sets_before = 0;
};

extern "C" {

int fgs_gmt_parse(const char *text, size_t len, const char *const *gene_names, int n_gene,
                  int gs_size_min, int gs_size_max, fgs_gmt **out) {
  if (!text || !out || (n_gene > 0 && !gene_names) || n_gene < 0) {
    fgs::set_error("fgs_gmt_parse: bad argument");
    return FGS_ERR_ARG;
  }
  std::unordered_map<std::string_view, int> first;  // gene name -> first 1-based position
  first.reserve((size_t)n_gene * 2);
  for (int i = 0; i < n_gene; i++) {
    if (!gene_names[i]) { fgs::set_error("fgs_gmt_parse: gene name %d is NULL", i); return FGS_ERR_ARG; }
    first.emplace(std::string_view(gene_names[i]), i + 1);  // emplace keeps the first
  }
  fgs_gmt *g = new fgs_gmt();
  g->offsets.push_back(0);
  std::vector<int> cur;
  size_t pos = 0;
  while (pos < len) {
    size_t eol = pos;
    while (eol < len && text[eol] != '\n') eol++;
    size_t end = eol;
    if (end > pos && text[end - 1] == '\r') end--;
    if (end > pos) {  // empty lines are skipped
      // field 1 = name, field 2 = description, the rest = genes
      size_t t1 = pos;
      while (t1 < end && text[t1] != '\t') t1++;
      size_t t2 = t1 < end ? t1 + 1 : end;
      while (t2 < end && text[t2] != '\t') t2++;
      g->sets_before++;
      cur.clear();
      size_t p = t2 < end ? t2 + 1 : end;
      bool any = t2 < end;
      while (any) {
        size_t q = p;
        while (q < end && text[q] != '\t') q++;
        g->genes_before++;
        auto it = first.find(std::string_view(text + p, q - p));
        if (it != first.end()) cur.push_back(it->second);
        if (q >= end) break;
        p = q + 1;
      }
      g->genes_after += (long long)cur.size();
      if ((int)cur.size() >= gs_size_min && (int)cur.size() <= gs_size_max) {
        g->names.emplace_back(text + pos, t1 - pos);
        g->index.insert(g->index.end(), cur.begin(), cur.end());
        if (g->index.size() > (size_t)0x7fffffff) {
          delete g;
          fgs::set_error("fgs_gmt_parse: more than 2^31 - 1 set members");
          return FGS_ERR_ARG;
        }
        g->offsets.push_back((int)g->index.size());
      }
    }
    pos = eol + 1;
  }
  *out = g;
  return FGS_OK;
}

int fgs_gmt_read(const char *path, const char *const *gene_names, int n_gene, int gs_size_min,
                 int gs_size_max, fgs_gmt **out) {
  if (!path) { fgs::set_error("fgs_gmt_read: path is NULL"); return FGS_ERR_ARG; }
  FILE *f = fopen(path, "rb");
  if (!f) { fgs::set_error("fgs_gmt_read: cannot open '%s'", path); return FGS_ERR_ARG; }
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
  fclose(f);
  return fgs_gmt_parse(text.data(), text.size(), gene_names, n_gene, gs_size_min, gs_size_max, out);
}

int fgs_gmt_n_sets(const fgs_gmt *g) { return g ? (int)g->names.size() : 0; }
long long fgs_gmt_nnz(const fgs_gmt *g) { return g ? (long long)g->index.size() : 0; }
const int *fgs_gmt_offsets(const fgs_gmt *g) { return g ? g->offsets.data() : nullptr; }
const int *fgs_gmt_index(const fgs_gmt *g) { return g ? g->index.data() : nullptr; }
const char *fgs_gmt_set_name(const fgs_gmt *g, int k) {
  return (g && k >= 0 && k < (int)g->names.size()) ? g->names[k].c_str() : nullptr;
}
int fgs_gmt_stats(const fgs_gmt *g, long long *genes_before, long long *genes_after, int *sets_before) {
  if (!g) return FGS_ERR_ARG;
  if (genes_before) *genes_before = g->genes_before;
  if (genes_after) *genes_after = g->genes_after;
  if (sets_before) *sets_before = g->sets_before;
  return FGS_OK;
}
void fgs_gmt_free(fgs_gmt *g) { delete g; }

}  // extern "C"
