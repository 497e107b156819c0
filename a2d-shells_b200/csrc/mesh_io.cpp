// mesh_io.cpp — the data format in front of the path: NASTRAN bulk-data (.bdf) decks and a
// binary mesh container, behind the C ABI of include/a2ds.h (a2ds_mesh_*).  Host only.
//
// Stands in for TACSMeshLoader::scanBDFFile / getConnectivity / getBCs
// (src/io/TACSMeshLoader.cpp:570-1096, 1221-1278): same cards (GRID, GRID*, SPC, SPC*, the
// ten element keywords of src/io/TACSMeshLoader.h:20-35), same field columns, same node
// reordering per element type, same numbering of the result (nodes and elements sorted by
// their file numbers, 0-based), so the arrays are interchangeable with the reference
// loader's.  Checked array by array against it in tests/test_mesh_io.py.
//
// Built differently: the bulk section is cut into chunks at card boundaries and the chunks
// are parsed concurrently (one pass each, growing vectors — the reference walks the file
// twice on one thread), numbers are converted in place without a format string, the result
// is merged in chunk order so that it does not depend on the thread count.  A parsed mesh
// can be written to / read from a flat binary file (no parsing at all on the next start:
// what a 16 M-element run wants).
//
// Deliberate differences to the reference reader, all where it reads memory it did not
// write:
//   - columns beyond the end of a line are blank (the reference re-reads bytes left in its
//     line buffer by the previous, longer line);
//   - only the first 80 columns of a line count (its 81-byte buffer is not terminated when
//     a line is longer);
//   - nodes 9..27 of a CHEXA27 are taken in file order (the reference leaves them
//     uninitialised, src/io/TACSMeshLoader.cpp:945-953), and a CQUAD8 keeps its file order (the
//     reference applies the 9-node permutation to it and reads a ninth node the card does not
//     have, :938-948).
// And one where it is silent: an element card without its numbers, or with fewer nodes than
// its keyword needs, ends the reference's scan without an error (the failure flag set at
// :842 / :914 is a shadowed local) and a truncated mesh comes back; here the call fails.
// A2DS_MESH_TIMING=1 in the environment prints the time of each stage to stderr.
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/a2ds.h"

int a2ds_set_error_(const char *msg);  // a2ds.cu: stores the message for a2ds_last_error()

namespace {

int failm(const std::string &m) { return a2ds_set_error_(m.c_str()); }

// no exception leaves the C ABI: out-of-memory and the like become an error return
template <class F>
int guarded(const char *what, F &&body) {
  try {
    return body();
  } catch (const std::exception &e) {
    return failm(std::string(what) + ": " + e.what());
  }
}

// element keywords in matching order (longer keywords before their prefixes) with the
// admissible node counts — src/io/TACSMeshLoader.h:20-35
struct ElemKind { const char *key; int len, nmin, nmax; };
const ElemKind KINDS[] = {{"CBAR", 4, 2, 2},     {"CQUADR", 6, 4, 4},  {"CQUAD4", 6, 4, 4},
                          {"CQUAD8", 6, 8, 8},   {"CQUAD9", 6, 9, 9},  {"CQUAD", 5, 9, 9},
                          {"CHEXA27", 7, 27, 27}, {"CHEXA", 5, 8, 8},  {"CTRIA3", 6, 3, 3},
                          {"CTETRA", 6, 4, 10}};
const int N_KINDS = sizeof(KINDS) / sizeof(KINDS[0]);

// file order -> tensor-product order of the element basis (src/io/TACSMeshLoader.cpp:932-957)
const int ORDER_QUAD4[4] = {0, 1, 3, 2};
const int ORDER_QUAD9[9] = {0, 4, 1, 7, 8, 5, 3, 6, 2};
const int ORDER_HEXA[8] = {0, 1, 3, 2, 4, 5, 7, 6};

struct Line {
  const char *p;
  int len;  // columns that count (<= 80), without the newline
  bool starts(const char *s, int n) const { return len >= n && memcmp(p, s, n) == 0; }
};

// one line starting at `at`; returns the offset of the next line
size_t take_line(const char *buf, size_t at, size_t end, Line &ln) {
  const char *nl = (const char *)memchr(buf + at, '\n', end - at);
  const size_t stop = nl ? (size_t)(nl - buf) : end;
  ln.p = buf + at;
  ln.len = (int)std::min<size_t>(stop - at, 80);
  return stop + 1;
}

// columns [col, col + width) of a line as a terminated string; returns how many of them exist
int field(const Line &ln, int col, int width, char *out) {
  int n = ln.len - col;
  n = n < 0 ? 0 : (n > width ? width : n);
  memcpy(out, ln.p + col, n);
  out[n] = '\0';
  return n;
}

int to_int(const char *s) { return (int)strtol(s, nullptr, 10); }

// strtod for the strings a deck holds.  A decimal with at most 15 significant digits and a
// power of ten within 10^+-22 is the correctly rounded quotient / product of two exactly
// representable doubles, so one multiplication or division gives the same bits strtod does;
// anything else (longer mantissa, large exponent, inf / nan / hex, malformed text) goes to
// strtod itself.  ~10x faster on GRID coordinates, and identical by construction.
double decimal(const char *s) {
  static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                                 1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char *c = s;
  while (isspace((unsigned char)*c)) c++;
  bool neg = false;
  if (*c == '-' || *c == '+') neg = (*c++ == '-');
  uint64_t mant = 0;
  int digits = 0, seen = 0, exp10 = 0;
  for (; *c >= '0' && *c <= '9'; c++, seen++) {
    if (mant || *c != '0') digits++;
    if (digits > 15) return strtod(s, nullptr);
    mant = 10 * mant + (uint64_t)(*c - '0');
  }
  if (*c == '.') {
    for (c++; *c >= '0' && *c <= '9'; c++, seen++) {
      if (mant || *c != '0') digits++;
      if (digits > 15) return strtod(s, nullptr);
      mant = 10 * mant + (uint64_t)(*c - '0');
      exp10--;
    }
  }
  if (!seen) return strtod(s, nullptr);  // no number here (or inf / nan): let strtod decide
  if (*c == 'e' || *c == 'E') {
    const char *e = c + 1;
    bool eneg = false;
    if (*e == '-' || *e == '+') eneg = (*e++ == '-');
    if (*e >= '0' && *e <= '9') {  // otherwise the letter is not part of the number
      int ex = 0;
      for (; *e >= '0' && *e <= '9'; e++)
        if (ex < 10000) ex = 10 * ex + (*e - '0');
      exp10 += eneg ? -ex : ex;
      c = e;
    }
  }
  if (*c == 'x' || *c == 'X' || *c == 'p' || *c == 'P') return strtod(s, nullptr);  // hex float
  double v;
  if (mant == 0) v = 0.0;
  else if (exp10 == 0) v = (double)mant;
  else if (exp10 > 0 && exp10 <= 22) v = (double)mant * P10[exp10];
  else if (exp10 < 0 && exp10 >= -22) v = (double)mant / P10[-exp10];
  else return strtod(s, nullptr);
  return neg ? -v : v;
}

// NASTRAN reals: "1.5E-3", "1.5D-3" and the compact "1.5-3" (exponent sign without a letter).
// Same outcomes as bdf_atof (src/io/TACSMeshLoader.cpp:177-219), including its corner
// cases: blank -> 0, and a sign anywhere after the first character opens the exponent.
double to_real(const char *s) {
  char t[48];
  int n = 0;
  bool letter = false;
  for (const char *c = s; *c; c++)
    if (*c == 'e' || *c == 'E' || *c == 'd' || *c == 'D') { letter = true; break; }
  if (letter) {
    for (const char *c = s; *c && n < 46; c++) {
      // only the first exponent letter is looked at by the reference; a 'd' becomes 'e'
      t[n++] = (*c == 'd' || *c == 'D') ? 'e' : *c;
      if (*c == 'd' || *c == 'D' || *c == 'e' || *c == 'E') {
        for (c++; *c && n < 46; c++) t[n++] = *c;
        break;
      }
    }
    t[n] = '\0';
    return decimal(t);
  }
  const char *c = s;
  while (*c == ' ') c++;
  if (!*c) return 0.0;
  if (*c == '-') t[n++] = *c++;
  for (; *c && n < 46; c++) {
    if (*c == '-' || *c == '+') t[n++] = 'e';
    t[n++] = *c;
  }
  t[n] = '\0';
  return decimal(t);
}

// what one chunk of the bulk section contains, in file order
struct Parsed {
  std::vector<int> node_id;
  std::vector<double> xyz;
  std::vector<int> elem_id, elem_comp, elem_kind, conn, conn_ptr{0};
  std::vector<int> bc_node, bc_ptr{0}, bc_var;
  std::vector<double> bc_val;
  std::vector<std::string> shell_names;  // "$       Shell" comment lines (ICEM component names)
  std::vector<std::string> unknown;      // first few unrecognised cards
  long n_unknown = 0;
  bool ended = false;                    // ENDDATA / END BULK seen
  std::string error;                     // hard failure: nothing after it is read
};

void parse_spc(const Line &ln, int node_col, int dof_col, int val_col, int w, Parsed &out) {
  char f[20];
  field(ln, node_col, w, f);
  out.bc_node.push_back(to_int(f) - 1);
  field(ln, val_col, w, f);
  const double val = to_real(f);
  for (int k = dof_col; k < dof_col + w && k < ln.len; k++) {
    const char ch = ln.p[k];
    if (ch >= '1' && ch <= '8') {
      out.bc_var.push_back(ch - '1');
      out.bc_val.push_back(val);
    }
  }
  out.bc_ptr.push_back((int)out.bc_var.size());
}

void parse_grid(const Line &ln, Parsed &out) {
  char f[5][32] = {{0}, {0}, {0}, {0}, {0}};
  const bool comma = memchr(ln.p, ',', ln.len) != nullptr;
  if (comma) {
    // keyword in the first 8 columns, then comma separated fields
    int end = 8;
    for (int i = 0; i < 5; i++) {
      const int start = end;
      while (end < ln.len && ln.p[end] != ',') end++;
      int n = end - start;
      n = n < 0 ? 0 : (n > 31 ? 31 : n);  // the reference's field buffers hold 31 characters
      if (start < ln.len) memcpy(f[i], ln.p + start, n);
      f[i][n] = '\0';
      end++;
    }
  } else {
    field(ln, 8, 8, f[0]);
    field(ln, 24, 8, f[2]);
    field(ln, 32, 8, f[3]);
    field(ln, 40, 8, f[4]);
  }
  out.node_id.push_back(to_int(f[0]) - 1);
  out.xyz.push_back(to_real(f[2]));
  out.xyz.push_back(to_real(f[3]));
  out.xyz.push_back(to_real(f[4]));
}

// An element card and its continuation lines.  Fields are `w` columns wide from column 8:
// element number, component number, then node numbers until the keyword's maximum is reached
// or a non-positive entry ends the list; blank fields are skipped; a continuation line starts
// with '*' or ' ' (src/io/TACSMeshLoader.cpp:319-421).  Returns the offset of the next card,
// or 0 after a hard failure.
size_t parse_element(const char *buf, size_t at, size_t end, int kind, int w, Parsed &out) {
  Line ln;
  size_t next = take_line(buf, at, end, ln);
  char f[20];
  field(ln, 8, w, f);
  const int eid = to_int(f);
  field(ln, 8 + w, w, f);
  const int comp = to_int(f);
  if (ln.len <= 0 || eid <= 0 || comp <= 0) {
    out.error = "element card without a positive element / component number: " +
                std::string(ln.p, ln.len);
    return 0;
  }
  const int nmax = KINDS[kind].nmax;
  int nodes[27], n = 0, col = 8 + 2 * w;
  bool open = true;
  while (open && n < nmax) {
    for (; n < nmax && col < ln.len; col += w) {
      const int have = field(ln, col, w, f);
      bool blank = (have == w);  // a field cut short by the end of the line is not "blank"
      for (int k = 0; blank && k < have; k++) blank = isspace((unsigned char)f[k]) != 0;
      if (blank) continue;
      const int v = to_int(f);
      if (v <= 0) { open = false; break; }
      nodes[n++] = v;
    }
    if (!open || n >= nmax) break;
    if (next >= end) break;
    Line cont;
    const size_t after = take_line(buf, next, end, cont);
    if (cont.len <= 0 || !(cont.p[0] == '*' || cont.p[0] == ' ')) break;
    ln = cont;
    next = after;
    col = 8;
  }
  if (n < KINDS[kind].nmin) {
    out.error = std::string("number of nodes for element ") + KINDS[kind].key + " not within limits";
    return 0;
  }
  const int *order = nullptr;
  int n_order = 0;
  const char *key = KINDS[kind].key;
  if (!strcmp(key, "CQUAD4") || !strcmp(key, "CQUADR")) { order = ORDER_QUAD4; n_order = 4; }
  else if (!strcmp(key, "CQUAD9") || !strcmp(key, "CQUAD")) { order = ORDER_QUAD9; n_order = 9; }
  else if (!strcmp(key, "CHEXA") || !strcmp(key, "CHEXA27")) { order = ORDER_HEXA; n_order = 8; }
  for (int k = 0; k < n; k++) out.conn.push_back(nodes[(k < n_order) ? order[k] : k] - 1);
  out.conn_ptr.push_back((int)out.conn.size());
  out.elem_id.push_back(eid - 1);
  out.elem_comp.push_back(comp - 1);
  out.elem_kind.push_back((kind == 9 && n == 10) ? N_KINDS : kind);  // N_KINDS: "CTETRA10"
  return next;
}

void parse_chunk(const char *buf, size_t at, size_t end, Parsed &out) {
  Line ln;
  while (at < end) {
    size_t next = take_line(buf, at, end, ln);
    if (ln.len == 0) {  // the reference stops with a failure at an empty line (:772-776)
      out.error = "empty line inside the bulk data";
      return;
    }
    if (ln.starts("$       Shell", 13)) {
      // ICEM writes one such comment per component: its name sits in columns 41..72
      char name[40];
      const int have = ln.len > 41 ? std::min(ln.len - 41, 32) : 0;
      memcpy(name, ln.p + 41, have);
      name[have] = '\0';
      char tok[40] = "";
      sscanf(name, "%32s", tok);
      out.shell_names.push_back(tok);
    }
    if (ln.p[0] != '$') {
      if (ln.starts("END BULK", 8) || ln.starts("ENDDATA", 7)) {
        out.ended = true;
        return;
      } else if (ln.starts("GRID*", 5)) {
        // large field: number in columns 8..24, x 40..56, y 56..72, z on the next line 8..24
        Line l2;
        if (next >= end) { out.error = "GRID* without its second line"; return; }
        next = take_line(buf, next, end, l2);
        if (l2.len == 0) { out.error = "GRID* without its second line"; return; }
        char f[20];
        field(ln, 8, 16, f);  out.node_id.push_back(to_int(f) - 1);
        field(ln, 40, 16, f); out.xyz.push_back(to_real(f));
        field(ln, 56, 16, f); out.xyz.push_back(to_real(f));
        field(l2, 8, 16, f);  out.xyz.push_back(to_real(f));
      } else if (ln.starts("GRID", 4)) {
        parse_grid(ln, out);
      } else if (ln.starts("SPC*", 4)) {
        parse_spc(ln, 24, 40, 56, 16, out);
      } else if (ln.starts("SPC", 3)) {
        parse_spc(ln, 16, 24, 32, 8, out);
      } else {
        int kind = -1;
        for (int k = 0; k < N_KINDS && kind < 0; k++)
          if (ln.starts(KINDS[k].key, KINDS[k].len)) kind = k;
        if (kind >= 0) {
          const int w = (ln.len > KINDS[kind].len && ln.p[KINDS[kind].len] == '*') ? 16 : 8;
          next = parse_element(buf, at, end, kind, w, out);
          if (!next) return;
        } else {
          if (out.unknown.size() < 4) out.unknown.emplace_back(ln.p, ln.len);
          out.n_unknown++;
        }
      }
    }
    at = next;
  }
}

// offset after the last line that starts with "BEGIN BULK" (0: the whole file is bulk data,
// :741-757)
size_t find_bulk_start(const char *buf, size_t len) {
  size_t start = 0;
  const char *p = buf, *stop = buf + len;
  while ((p = (const char *)memmem(p, stop - p, "BEGIN BULK", 10)) != nullptr) {
    if (p == buf || p[-1] == '\n') {
      const char *nl = (const char *)memchr(p, '\n', stop - p);
      if (nl && (size_t)(nl + 1 - buf) < len) start = (size_t)(nl + 1 - buf);
    }
    p += 10;
  }
  return start;
}

// chunk boundaries: the start of a line that opens a card (a letter or '$' in column 0) and
// is not the second line of a GRID* entry
std::vector<size_t> cut_points(const char *buf, size_t begin, size_t end, int parts) {
  std::vector<size_t> cuts{begin};
  for (int k = 1; k < parts; k++) {
    size_t at = begin + (end - begin) / parts * k;
    if (at <= cuts.back()) continue;
    // back to a line start, remembering the line before it
    const char *nl = (const char *)memchr(buf + at, '\n', end - at);
    if (!nl) break;
    size_t prev = at;  // start of the line that contains `at`
    while (prev > begin && buf[prev - 1] != '\n') prev--;
    at = (size_t)(nl - buf) + 1;
    while (at < end) {
      const char c = buf[at];
      const bool opens = (c >= 'A' && c <= 'Z') || c == '$';
      const bool after_grid_star = (end - prev >= 5) && memcmp(buf + prev, "GRID*", 5) == 0;
      if (opens && !after_grid_star) break;
      prev = at;
      const char *n2 = (const char *)memchr(buf + at, '\n', end - at);
      if (!n2) { at = end; break; }
      at = (size_t)(n2 - buf) + 1;
    }
    if (at < end && at > cuts.back()) cuts.push_back(at);
  }
  cuts.push_back(end);
  return cuts;
}

}  // namespace

struct a2ds_mesh {
  std::vector<int> elem_ptr, elem_conn, elem_comp;
  std::vector<double> X;
  std::vector<int> bc_nodes, bc_ptr, bc_vars;
  std::vector<double> bc_vals;
  std::vector<int> node_nums, elem_nums;  // file numbers (0-based), ascending
  int n_comp = 0;
  std::vector<char> comp_elem, comp_name;  // 9 / 33 characters per component
  long n_unknown = 0;
};

static int read_bdf_impl(const char *path, int n_threads, a2ds_mesh **out) {
  *out = nullptr;
  FILE *fp = fopen(path, "rb");
  if (!fp) return failm(std::string("a2ds_mesh_read_bdf: unable to open file ") + path);
  fseek(fp, 0, SEEK_END);
  const size_t len = (size_t)ftell(fp);
  rewind(fp);
  std::vector<char> data(len + 1);
  if (len && fread(data.data(), 1, len, fp) != len) {
    fclose(fp);
    return failm(std::string("a2ds_mesh_read_bdf: problem reading file ") + path);
  }
  fclose(fp);
  const char *buf = data.data();
  auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (getenv("A2DS_MESH_TIMING")) {
      auto T1 = std::chrono::steady_clock::now();
      fprintf(stderr, "  [mesh_io] %-12s %.3f s\n", what, std::chrono::duration<double>(T1 - T0).count());
      T0 = T1;
    }
  };
  lap("read");
  const size_t begin = find_bulk_start(buf, len);
  lap("bulk start");

  if (n_threads <= 0) n_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (len - begin < (1u << 20)) n_threads = 1;  // not worth a thread below a megabyte
  const std::vector<size_t> cuts = cut_points(buf, begin, len, n_threads);
  const int parts = (int)cuts.size() - 1;
  std::vector<Parsed> chunk(parts);
  if (parts == 1) {
    parse_chunk(buf, cuts[0], cuts[1], chunk[0]);
  } else {
    std::vector<std::thread> pool;
    for (int k = 0; k < parts; k++)
      pool.emplace_back([&, k] {
        try {
          parse_chunk(buf, cuts[k], cuts[k + 1], chunk[k]);
        } catch (const std::exception &e) {  // out of memory in a worker: report, do not terminate
          chunk[k].error = std::string("parser thread: ") + e.what();
        }
      });
    for (auto &t : pool) t.join();
  }

  if (getenv("A2DS_MESH_TIMING")) {
    fprintf(stderr, "  [mesh_io] %d chunk(s):", parts);
    for (int k = 0; k < parts; k++) fprintf(stderr, " %zu", cuts[k + 1] - cuts[k]);
    fprintf(stderr, "\n");
  }
  lap("parse");
  // merge in file order; a chunk that ended the deck (ENDDATA) hides everything behind it
  Parsed all;
  for (int k = 0; k < parts; k++) {
    Parsed &c = chunk[k];
    const int conn0 = (int)all.conn.size(), var0 = (int)all.bc_var.size();
    all.node_id.insert(all.node_id.end(), c.node_id.begin(), c.node_id.end());
    all.xyz.insert(all.xyz.end(), c.xyz.begin(), c.xyz.end());
    all.elem_id.insert(all.elem_id.end(), c.elem_id.begin(), c.elem_id.end());
    all.elem_comp.insert(all.elem_comp.end(), c.elem_comp.begin(), c.elem_comp.end());
    all.elem_kind.insert(all.elem_kind.end(), c.elem_kind.begin(), c.elem_kind.end());
    all.conn.insert(all.conn.end(), c.conn.begin(), c.conn.end());
    for (size_t i = 1; i < c.conn_ptr.size(); i++) all.conn_ptr.push_back(conn0 + c.conn_ptr[i]);
    all.bc_node.insert(all.bc_node.end(), c.bc_node.begin(), c.bc_node.end());
    all.bc_var.insert(all.bc_var.end(), c.bc_var.begin(), c.bc_var.end());
    all.bc_val.insert(all.bc_val.end(), c.bc_val.begin(), c.bc_val.end());
    for (size_t i = 1; i < c.bc_ptr.size(); i++) all.bc_ptr.push_back(var0 + c.bc_ptr[i]);
    all.shell_names.insert(all.shell_names.end(), c.shell_names.begin(), c.shell_names.end());
    for (auto &u : c.unknown)
      if (all.unknown.size() < 4) all.unknown.push_back(u);
    all.n_unknown += c.n_unknown;
    if (!c.error.empty()) return failm("a2ds_mesh_read_bdf: " + c.error);
    if (c.ended) break;
    c = Parsed();  // release the chunk's memory as we go
  }
  for (auto &u : all.unknown)  // as the reference: reported, not fatal (:861-865)
    fprintf(stderr, "a2ds_mesh_read_bdf: card not recognized. Line\n %s\n", u.c_str());
  if (all.n_unknown > (long)all.unknown.size())
    fprintf(stderr, "a2ds_mesh_read_bdf: ... and %ld more\n", all.n_unknown - (long)all.unknown.size());

  lap("merge");
  const int nn = (int)all.node_id.size(), ne = (int)all.elem_id.size(), nb = (int)all.bc_node.size();
  a2ds_mesh *m = new a2ds_mesh();
  m->n_unknown = all.n_unknown;
  // nodes and elements in ascending file number (stable: ties keep file order)
  std::vector<int> nperm(nn), eperm(ne);
  std::iota(nperm.begin(), nperm.end(), 0);
  std::iota(eperm.begin(), eperm.end(), 0);
  if (!std::is_sorted(all.node_id.begin(), all.node_id.end()))
    std::stable_sort(nperm.begin(), nperm.end(),
                     [&](int a, int b) { return all.node_id[a] < all.node_id[b]; });
  if (!std::is_sorted(all.elem_id.begin(), all.elem_id.end()))
    std::stable_sort(eperm.begin(), eperm.end(),
                     [&](int a, int b) { return all.elem_id[a] < all.elem_id[b]; });
  m->node_nums.resize(nn);
  m->X.resize(3 * (size_t)nn);
  for (int k = 0; k < nn; k++) {
    m->node_nums[k] = all.node_id[nperm[k]];
    for (int j = 0; j < 3; j++) m->X[3 * (size_t)k + j] = all.xyz[3 * (size_t)nperm[k] + j];
  }
  // file number -> node: a direct table when the numbers are dense, else a binary search
  std::vector<int> table;
  const long span = nn ? (long)m->node_nums.back() - (long)m->node_nums.front() + 1 : 0;
  const int first_num = nn ? m->node_nums.front() : 0;
  if (nn && span <= 4L * nn + 1024) {
    table.assign((size_t)span, -1);
    for (int k = nn - 1; k >= 0; k--) table[m->node_nums[k] - first_num] = k;  // ties: first one
  }
  auto lookup = [&](int file_num) {
    if (!table.empty()) {
      const long o = (long)file_num - first_num;
      return (o >= 0 && o < span) ? table[(size_t)o] : -1;
    }
    auto it = std::lower_bound(m->node_nums.begin(), m->node_nums.end(), file_num);
    return (it != m->node_nums.end() && *it == file_num) ? (int)(it - m->node_nums.begin()) : -1;
  };
  int missing = 0, missing_num = 0;
  for (int c : all.elem_comp) m->n_comp = std::max(m->n_comp, c + 1);
  m->elem_nums.resize(ne);
  m->elem_comp.resize(ne);
  m->elem_ptr.assign(1, 0);
  m->elem_conn.reserve(all.conn.size());
  for (int k = 0; k < ne; k++) {
    const int e = eperm[k];
    m->elem_nums[k] = all.elem_id[e];
    m->elem_comp[k] = all.elem_comp[e];
    for (int j = all.conn_ptr[e]; j < all.conn_ptr[e + 1]; j++) {
      const int node = lookup(all.conn[j]);
      if (node < 0 && !missing++) missing_num = all.conn[j] + 1;
      m->elem_conn.push_back(node);
    }
    m->elem_ptr.push_back((int)m->elem_conn.size());
  }
  m->bc_nodes.resize(nb);
  for (int k = 0; k < nb; k++) {
    m->bc_nodes[k] = lookup(all.bc_node[k]);
    if (m->bc_nodes[k] < 0 && !missing++) missing_num = all.bc_node[k] + 1;
  }
  m->bc_ptr = all.bc_ptr;
  m->bc_vars = all.bc_var;
  m->bc_vals = all.bc_val;
  // per component: keyword of the first element that uses it (file order), ICEM name
  m->comp_elem.assign(9 * (size_t)m->n_comp, '\0');
  m->comp_name.assign(33 * (size_t)m->n_comp, '\0');
  for (int e = 0; e < ne; e++) {
    char *dst = &m->comp_elem[9 * (size_t)all.elem_comp[e]];
    if (!dst[0]) {
      const int kd = all.elem_kind[e];
      snprintf(dst, 9, "%s", kd == N_KINDS ? "CTETRA10" : KINDS[kd].key);
    }
  }
  for (size_t k = 0; k < all.shell_names.size() && (int)k < m->n_comp; k++)
    snprintf(&m->comp_name[33 * k], 33, "%s", all.shell_names[k].c_str());
  if (missing) {
    delete m;
    return failm("a2ds_mesh_read_bdf: " + std::to_string(missing) +
                 " reference(s) to undefined grid points, first: " + std::to_string(missing_num));
  }
  lap("number");
  *out = m;
  return 0;
}

extern "C" void a2ds_mesh_free(a2ds_mesh *m) { delete m; }

extern "C" int a2ds_mesh_sizes(const a2ds_mesh *m, int *n_nodes, int *n_elems, int *conn_size,
                               int *n_bcs, int *bc_size, int *n_comp) {
  if (!m) return failm("a2ds_mesh_sizes: null mesh");
  if (n_nodes) *n_nodes = (int)m->node_nums.size();
  if (n_elems) *n_elems = (int)m->elem_comp.size();
  if (conn_size) *conn_size = (int)m->elem_conn.size();
  if (n_bcs) *n_bcs = (int)m->bc_nodes.size();
  if (bc_size) *bc_size = (int)m->bc_vars.size();
  if (n_comp) *n_comp = m->n_comp;
  return 0;
}

extern "C" int a2ds_mesh_connectivity(const a2ds_mesh *m, const int **elem_ptr,
                                      const int **elem_conn, const int **elem_comp,
                                      const double **X) {
  if (!m) return failm("a2ds_mesh_connectivity: null mesh");
  if (elem_ptr) *elem_ptr = m->elem_ptr.data();
  if (elem_conn) *elem_conn = m->elem_conn.data();
  if (elem_comp) *elem_comp = m->elem_comp.data();
  if (X) *X = m->X.data();
  return 0;
}

extern "C" int a2ds_mesh_bcs(const a2ds_mesh *m, const int **bc_nodes, const int **bc_ptr,
                             const int **bc_vars, const double **bc_vals) {
  if (!m) return failm("a2ds_mesh_bcs: null mesh");
  if (bc_nodes) *bc_nodes = m->bc_nodes.data();
  if (bc_ptr) *bc_ptr = m->bc_ptr.data();
  if (bc_vars) *bc_vars = m->bc_vars.data();
  if (bc_vals) *bc_vals = m->bc_vals.data();
  return 0;
}

extern "C" int a2ds_mesh_file_numbers(const a2ds_mesh *m, const int **node_nums,
                                      const int **elem_nums) {
  if (!m) return failm("a2ds_mesh_file_numbers: null mesh");
  if (node_nums) *node_nums = m->node_nums.data();
  if (elem_nums) *elem_nums = m->elem_nums.data();
  return 0;
}

extern "C" int a2ds_mesh_component(const a2ds_mesh *m, int comp, const char **elem_descript,
                                   const char **comp_descript) {
  if (!m || comp < 0 || comp >= m->n_comp) return failm("a2ds_mesh_component: no such component");
  if (elem_descript) *elem_descript = &m->comp_elem[9 * (size_t)comp];
  if (comp_descript) *comp_descript = &m->comp_name[33 * (size_t)comp];
  return 0;
}

// The arrays a2ds_set_mesh / a2ds_set_bcs take, for a deck of 4-node shells: connectivity
// 4 per element, one bit mask and six values per SPC card (DOF 7, 8 of a card are dropped:
// the shell has six).
extern "C" int a2ds_mesh_quad4(const a2ds_mesh *m, int *conn4, int *bc_masks, double *bc_vals6) {
  if (!m) return failm("a2ds_mesh_quad4: null mesh");
  const size_t ne = m->elem_comp.size();
  for (size_t e = 0; e < ne; e++)
    if (m->elem_ptr[e + 1] - m->elem_ptr[e] != 4)
      return failm("a2ds_mesh_quad4: element " + std::to_string(m->elem_nums[e] + 1) +
                   " does not have 4 nodes");
  if (conn4) memcpy(conn4, m->elem_conn.data(), 4 * ne * sizeof(int));
  for (size_t b = 0; b < m->bc_nodes.size(); b++) {
    int mask = 0;
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int j = m->bc_ptr[b]; j < m->bc_ptr[b + 1]; j++)
      if (m->bc_vars[j] < 6) { mask |= 1 << m->bc_vars[j]; v[m->bc_vars[j]] = m->bc_vals[j]; }
    if (bc_masks) bc_masks[b] = mask;
    if (bc_vals6) memcpy(&bc_vals6[6 * b], v, sizeof(v));
  }
  return 0;
}

// ---- binary container -----------------------------------------------------------------
// "A2DSMSH1", six int64 counts, then the arrays of a2ds_mesh in a fixed order, each padded
// to 8 bytes.  Native byte order; written and read on the same kind of machine.
namespace {
const char MAGIC[8] = {'A', '2', 'D', 'S', 'M', 'S', 'H', '1'};

template <class T>
bool put(FILE *fp, const std::vector<T> &v) {
  const size_t bytes = v.size() * sizeof(T), pad = (8 - bytes % 8) % 8;
  const char zero[8] = {0};
  return (bytes == 0 || fwrite(v.data(), 1, bytes, fp) == bytes) && fwrite(zero, 1, pad, fp) == pad;
}
template <class T>
bool get(FILE *fp, std::vector<T> &v, size_t n) {
  v.resize(n);
  const size_t bytes = n * sizeof(T), pad = (8 - bytes % 8) % 8;
  char skip[8];
  return (bytes == 0 || fread(v.data(), 1, bytes, fp) == bytes) && fread(skip, 1, pad, fp) == pad;
}
}  // namespace

extern "C" int a2ds_mesh_write_bin(const a2ds_mesh *m, const char *path) {
  if (!m) return failm("a2ds_mesh_write_bin: null mesh");
  FILE *fp = fopen(path, "wb");
  if (!fp) return failm(std::string("a2ds_mesh_write_bin: unable to open file ") + path);
  const int64_t counts[6] = {(int64_t)m->node_nums.size(), (int64_t)m->elem_comp.size(),
                             (int64_t)m->elem_conn.size(), (int64_t)m->bc_nodes.size(),
                             (int64_t)m->bc_vars.size(), (int64_t)m->n_comp};
  bool ok = fwrite(MAGIC, 1, 8, fp) == 8 && fwrite(counts, sizeof(int64_t), 6, fp) == 6;
  ok = ok && put(fp, m->elem_ptr) && put(fp, m->elem_conn) && put(fp, m->elem_comp) &&
       put(fp, m->X) && put(fp, m->bc_nodes) && put(fp, m->bc_ptr) && put(fp, m->bc_vars) &&
       put(fp, m->bc_vals) && put(fp, m->node_nums) && put(fp, m->elem_nums) &&
       put(fp, m->comp_elem) && put(fp, m->comp_name);
  ok = (fclose(fp) == 0) && ok;
  return ok ? 0 : failm(std::string("a2ds_mesh_write_bin: short write to ") + path);
}

static int read_bin_impl(const char *path, a2ds_mesh **out) {
  *out = nullptr;
  FILE *fp = fopen(path, "rb");
  if (!fp) return failm(std::string("a2ds_mesh_read_bin: unable to open file ") + path);
  char magic[8];
  int64_t c[6];
  if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, MAGIC, 8) != 0 ||
      fread(c, sizeof(int64_t), 6, fp) != 6) {
    fclose(fp);
    return failm(std::string("a2ds_mesh_read_bin: not a mesh container: ") + path);
  }
  for (int k = 0; k < 6; k++)
    if (c[k] < 0 || c[k] > 0x7fffffff) {
      fclose(fp);
      return failm(std::string("a2ds_mesh_read_bin: corrupt header in ") + path);
    }
  // the header fixes the file size: check it before allocating anything from it
  {
    auto padded = [](int64_t n, int64_t size) { return (n * size + 7) / 8 * 8; };
    const int64_t want = 8 + 6 * 8 + padded(c[1] + 1, 4) + padded(c[2], 4) + padded(c[1], 4) +
                         padded(3 * c[0], 8) + padded(c[3], 4) + padded(c[3] + 1, 4) +
                         padded(c[4], 4) + padded(c[4], 8) + padded(c[0], 4) + padded(c[1], 4) +
                         padded(9 * c[5], 1) + padded(33 * c[5], 1);
    fseek(fp, 0, SEEK_END);
    const int64_t have = (int64_t)ftell(fp);
    fseek(fp, 8 + 6 * 8, SEEK_SET);
    if (have != want) {
      fclose(fp);
      return failm(std::string("a2ds_mesh_read_bin: truncated or inconsistent file ") + path);
    }
  }
  a2ds_mesh *m = new a2ds_mesh();
  m->n_comp = (int)c[5];
  bool ok = get(fp, m->elem_ptr, c[1] + 1) && get(fp, m->elem_conn, c[2]) &&
            get(fp, m->elem_comp, c[1]) && get(fp, m->X, 3 * c[0]) && get(fp, m->bc_nodes, c[3]) &&
            get(fp, m->bc_ptr, c[3] + 1) && get(fp, m->bc_vars, c[4]) && get(fp, m->bc_vals, c[4]) &&
            get(fp, m->node_nums, c[0]) && get(fp, m->elem_nums, c[1]) &&
            get(fp, m->comp_elem, 9 * c[5]) && get(fp, m->comp_name, 33 * c[5]);
  fclose(fp);
  // the arrays index each other: refuse a file whose contents do not fit its header
  ok = ok && m->elem_ptr.front() == 0 && m->elem_ptr.back() == c[2] && m->bc_ptr.front() == 0 &&
       m->bc_ptr.back() == c[4];
  for (size_t k = 0; ok && k + 1 < m->elem_ptr.size(); k++) ok = m->elem_ptr[k] <= m->elem_ptr[k + 1];
  for (size_t k = 0; ok && k + 1 < m->bc_ptr.size(); k++) ok = m->bc_ptr[k] <= m->bc_ptr[k + 1];
  for (size_t k = 0; ok && k < m->elem_conn.size(); k++) ok = m->elem_conn[k] >= 0 && m->elem_conn[k] < c[0];
  for (size_t k = 0; ok && k < m->bc_nodes.size(); k++) ok = m->bc_nodes[k] >= 0 && m->bc_nodes[k] < c[0];
  for (size_t k = 0; ok && k < m->elem_comp.size(); k++) ok = m->elem_comp[k] >= 0 && m->elem_comp[k] < c[5];
  if (!ok) {
    delete m;
    return failm(std::string("a2ds_mesh_read_bin: truncated or inconsistent file ") + path);
  }
  *out = m;
  return 0;
}

// a mesh container from arrays already in memory (generated meshes -> binary file)
static int from_arrays_impl(int n_nodes, int n_elems, const int *elem_ptr, const int *elem_conn,
                            const int *elem_comp, const double *X, int n_bcs, const int *bc_nodes,
                            const int *bc_ptr, const int *bc_vars, const double *bc_vals,
                            a2ds_mesh **out) {
  *out = nullptr;
  if (n_nodes < 0 || n_elems < 0 || n_bcs < 0) return failm("a2ds_mesh_from_arrays: bad sizes");
  a2ds_mesh *m = new a2ds_mesh();
  m->elem_ptr.assign(elem_ptr, elem_ptr + n_elems + 1);
  m->elem_conn.assign(elem_conn, elem_conn + elem_ptr[n_elems]);
  if (elem_comp) m->elem_comp.assign(elem_comp, elem_comp + n_elems);
  else m->elem_comp.assign(n_elems, 0);
  m->X.assign(X, X + 3 * (size_t)n_nodes);
  m->bc_ptr.assign(1, 0);
  if (n_bcs) {
    m->bc_nodes.assign(bc_nodes, bc_nodes + n_bcs);
    m->bc_ptr.assign(bc_ptr, bc_ptr + n_bcs + 1);
    m->bc_vars.assign(bc_vars, bc_vars + bc_ptr[n_bcs]);
    m->bc_vals.assign(bc_vals, bc_vals + bc_ptr[n_bcs]);
  }
  m->node_nums.resize(n_nodes);
  m->elem_nums.resize(n_elems);
  std::iota(m->node_nums.begin(), m->node_nums.end(), 0);
  std::iota(m->elem_nums.begin(), m->elem_nums.end(), 0);
  for (int c : m->elem_comp) m->n_comp = std::max(m->n_comp, c + 1);
  m->comp_elem.assign(9 * (size_t)m->n_comp, '\0');
  m->comp_name.assign(33 * (size_t)m->n_comp, '\0');
  for (int v : m->elem_conn)
    if (v < 0 || v >= n_nodes) {
      delete m;
      return failm("a2ds_mesh_from_arrays: connectivity refers to a node outside [0, n_nodes)");
    }
  *out = m;
  return 0;
}

extern "C" int a2ds_mesh_read_bdf(const char *path, int n_threads, a2ds_mesh **out) {
  return guarded("a2ds_mesh_read_bdf", [&] { return read_bdf_impl(path, n_threads, out); });
}
extern "C" int a2ds_mesh_read_bin(const char *path, a2ds_mesh **out) {
  return guarded("a2ds_mesh_read_bin", [&] { return read_bin_impl(path, out); });
}
extern "C" int a2ds_mesh_from_arrays(int n_nodes, int n_elems, const int *elem_ptr,
                                     const int *elem_conn, const int *elem_comp, const double *X,
                                     int n_bcs, const int *bc_nodes, const int *bc_ptr,
                                     const int *bc_vars, const double *bc_vals, a2ds_mesh **out) {
  return guarded("a2ds_mesh_from_arrays", [&] {
    return from_arrays_impl(n_nodes, n_elems, elem_ptr, elem_conn, elem_comp, X, n_bcs, bc_nodes,
                            bc_ptr, bc_vars, bc_vals, out);
  });
}
