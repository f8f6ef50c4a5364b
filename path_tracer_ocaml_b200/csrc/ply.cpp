// ply.cpp — the subset of the PLY format the reference's ply_format library reads, behind the C ABI.
//
// Role in the reference: Ply.of_bigstring (ply_format/src/ply.ml:340-352) = magic "ply\n", header lines up
// to "end_header" (format line, `element <name> <count>` each followed by its `property` lines,
// ply.ml:260-299), then the body: binary_little_endian only (ascii / big-endian are "to do" errors,
// ply.ml:346-349); an element is either all-atomic properties (fixed row width, ply.ml:208-218) or exactly ONE
// list property (ply.ml:220-248); anything mixed fails ("TO DO: parse mixed list/non-list element").
// ganesha consumes it as: element "vertex" -> columns x, y, z (Float or Double), and the list PROPERTY named
// "vertex_indices" -> rows, every row with exactly 3 indices (ganesha/bin/main.ml:50-60,182-185).
//
// Deliberate difference: the reference reads Short/Ushort values through the 1-byte accessors
// (ply.ml:104-105) although it advances 2 bytes; here they are read as the 2-byte integers they are.
// Files whose shorts fit in 7 bits parse identically.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene.hpp"

namespace ptb {
namespace {

enum Ty { CHAR, UCHAR, SHORT, USHORT, INT, UINT, FLOAT, DOUBLE, BAD };
Ty type_of(const std::string &s) {
  if (s == "char" || s == "int8") return CHAR;  // ply.ml:77-80 accepts the two aliases uint8 / int8
  if (s == "uchar" || s == "uint8") return UCHAR;
  if (s == "short") return SHORT;
  if (s == "ushort") return USHORT;
  if (s == "int") return INT;
  if (s == "uint") return UINT;
  if (s == "float") return FLOAT;
  if (s == "double") return DOUBLE;
  return BAD;
}
int size_of(Ty t) {
  switch (t) {
    case CHAR: case UCHAR: return 1;
    case SHORT: case USHORT: return 2;
    case INT: case UINT: case FLOAT: return 4;
    default: return 8;
  }
}
struct Prop {
  bool list = false;
  Ty type = BAD, len_type = BAD;
  std::string name;
};
struct Elem {
  std::string name;
  long long count = 0;
  std::vector<Prop> props;
};

long long read_int(const unsigned char *p, Ty t) {
  switch (t) {
    case CHAR: return (signed char)p[0];
    case UCHAR: return p[0];
    case SHORT: { int16_t v; std::memcpy(&v, p, 2); return v; }
    case USHORT: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case INT: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case UINT: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    default: return 0;
  }
}
double read_float(const unsigned char *p, Ty t) {
  if (t == FLOAT) { float v; std::memcpy(&v, p, 4); return v; }
  double v; std::memcpy(&v, p, 8); return v;
}
std::vector<std::string> split(const std::string &s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i <= s.size()) {  // String.split ~on:' ' keeps empty fields; a doubled space is a parse error there too
    size_t j = s.find(' ', i);
    if (j == std::string::npos) j = s.size();
    out.push_back(s.substr(i, j - i));
    i = j + 1;
  }
  return out;
}

int parse(const unsigned char *b, long long len, std::vector<float> *xyz, std::vector<int32_t> *faces) {
  if (len < 4) return fail(PTB_E_INVALID, "Could not read ply header (not enough bytes)");
  if (std::memcmp(b, "ply\n", 4) != 0) return fail(PTB_E_INVALID, "expected file to start with \"ply\\n\"");
  long long pos = 4;
  std::vector<std::string> lines;
  bool ended = false;
  while (pos < len) {
    const void *nl = std::memchr(b + pos, '\n', (size_t)(len - pos));
    if (!nl) break;
    long long e = (const unsigned char *)nl - b;
    std::string l((const char *)b + pos, (size_t)(e - pos));
    if (!l.empty() && l.back() == '\r') l.pop_back();
    pos = e + 1;
    if (l == "end_header") { ended = true; break; }
    lines.push_back(l);
  }
  if (!ended) return fail(PTB_E_INVALID, "missing \"end_header\" line");
  std::string format;
  for (const std::string &l : lines)
    if (l.rfind("format ", 0) == 0) {
      std::vector<std::string> w = split(l);
      if (w.size() != 3 || w[2] != "1.0") return fail(PTB_E_INVALID, "cannot parse format line: " + l);
      format = w[1];
      break;
    }
  if (format.empty()) return fail(PTB_E_INVALID, "header has no format line");
  if (format == "ascii" || format == "binary_big_endian")
    return fail(PTB_E_UNSUPPORTED, "to do: handle message format " + format);  // ply.ml:346-349
  if (format != "binary_little_endian") return fail(PTB_E_INVALID, "unrecognized format " + format);
  std::vector<Elem> elems;
  for (const std::string &l : lines) {
    const bool is_e = l.rfind("element ", 0) == 0, is_p = l.rfind("property ", 0) == 0;
    if (!is_e && !is_p) continue;  // comments, obj_info, ...
    std::vector<std::string> w = split(l);
    if (is_e) {
      if (w.size() != 3) return fail(PTB_E_INVALID, "expected element: " + l);
      Elem e;
      e.name = w[1];
      e.count = std::atoll(w[2].c_str());
      if (e.count < 0) return fail(PTB_E_INVALID, "negative element count: " + l);
      elems.push_back(e);
    } else {
      if (elems.empty()) return fail(PTB_E_INVALID, "expected element: " + l);
      Prop p;
      if (w.size() == 5 && w[1] == "list") {
        p.list = true, p.len_type = type_of(w[2]), p.type = type_of(w[3]), p.name = w[4];
        if (p.len_type == BAD || p.type == BAD) return fail(PTB_E_INVALID, "unrecognized type: " + l);
      } else if (w.size() == 3) {
        p.type = type_of(w[1]), p.name = w[2];
        if (p.type == BAD) return fail(PTB_E_INVALID, "unrecognized type: " + l);
      } else {
        return fail(PTB_E_INVALID, "cannot parse property: " + l);
      }
      elems.back().props.push_back(p);
    }
  }
  bool have_v = false, have_f = false;
  for (const Elem &e : elems) {
    const bool one_list = e.props.size() == 1 && e.props[0].list;
    bool atomic = true;
    for (const Prop &p : e.props) atomic = atomic && !p.list;
    if (one_list) {
      const Prop &p = e.props[0];
      if (p.len_type == FLOAT || p.len_type == DOUBLE || p.type == FLOAT || p.type == DOUBLE)
        return fail(PTB_E_INVALID, "expected integer type in list property " + p.name);  // int_accessor_exn
      const bool want = p.name == "vertex_indices";
      const int ls = size_of(p.len_type), es = size_of(p.type);
      // the header is untrusted: a list row takes at least its length field, so a count the rest of the file
      // cannot hold is rejected before anything is sized from it
      if (e.count > (len - pos) / std::max(ls, 1)) return fail(PTB_E_INVALID, "ply body is truncated (element " + e.name + ")");
      if (want) faces->reserve((size_t)e.count * 3);
      for (long long i = 0; i < e.count; ++i) {
        if (pos + ls > len) return fail(PTB_E_INVALID, "ply body is truncated (element " + e.name + ")");
        long long n = read_int(b + pos, p.len_type);
        pos += ls;
        if (n < 0 || n > (len - pos) / std::max(es, 1)) return fail(PTB_E_INVALID, "ply body is truncated (element " + e.name + ")");
        if (want) {
          if (n != 3) return fail(PTB_E_INVALID, "expected every face to have exactly 3 vertices");  // ganesha main.ml:182-185
          for (int k = 0; k < 3; ++k) faces->push_back((int32_t)read_int(b + pos + k * es, p.type));
        }
        pos += n * es;
      }
      have_f = have_f || want;
    } else if (atomic) {
      long long width = 0;
      int off[3] = {-1, -1, -1};
      Ty ty[3] = {BAD, BAD, BAD};
      for (const Prop &p : e.props) {
        for (int a = 0; a < 3; ++a)
          if (p.name == (a == 0 ? "x" : a == 1 ? "y" : "z")) off[a] = (int)width, ty[a] = p.type;
        width += size_of(p.type);
      }
      if (e.count > (len - pos) / std::max<long long>(width, 1))  // (division: width * count must not overflow)
        return fail(PTB_E_INVALID, "ply body is truncated (element " + e.name + ")");
      if (e.name == "vertex") {
        for (int a = 0; a < 3; ++a) {
          if (off[a] < 0) return fail(PTB_E_INVALID, "vertex element has no x/y/z property");
          if (ty[a] != FLOAT && ty[a] != DOUBLE) return fail(PTB_E_INVALID, "floats_exn: expected Floats");  // main.ml:45-48
        }
        xyz->resize((size_t)e.count * 3);
        for (long long i = 0; i < e.count; ++i)
          for (int a = 0; a < 3; ++a) (*xyz)[3 * i + a] = (float)read_float(b + pos + i * width + off[a], ty[a]);
        have_v = true;
      }
      pos += width * e.count;
    } else {
      return fail(PTB_E_UNSUPPORTED, "TO DO: parse mixed list/non-list element");  // ply.ml:246
    }
  }
  if (!have_v) return fail(PTB_E_INVALID, "ply has no \"vertex\" element");
  if (!have_f) return fail(PTB_E_INVALID, "ply has no \"vertex_indices\" list property");
  const long long nv = (long long)xyz->size() / 3;
  for (int32_t v : *faces)
    if (v < 0 || v >= nv) return fail(PTB_E_INVALID, "face index out of range");
  return PTB_OK;
}

int hand_over(std::vector<float> &xyz, std::vector<int32_t> &faces, float **oxyz, int64_t *nv, int32_t **ofaces, int64_t *nf) {
  *oxyz = (float *)std::malloc(std::max<size_t>(xyz.size(), 1) * sizeof(float));
  *ofaces = (int32_t *)std::malloc(std::max<size_t>(faces.size(), 1) * sizeof(int32_t));
  if (!*oxyz || !*ofaces) {
    std::free(*oxyz), std::free(*ofaces);
    *oxyz = nullptr, *ofaces = nullptr;
    return fail(PTB_E_NOMEM, "ply: out of memory");
  }
  if (!xyz.empty()) std::memcpy(*oxyz, xyz.data(), xyz.size() * sizeof(float));
  if (!faces.empty()) std::memcpy(*ofaces, faces.data(), faces.size() * sizeof(int32_t));
  *nv = (int64_t)xyz.size() / 3, *nf = (int64_t)faces.size() / 3;
  return PTB_OK;
}

}  // namespace
}  // namespace ptb

using namespace ptb;

extern "C" {

int ptb_ply_parse_mesh(const void *bytes, int64_t len, float **xyz, int64_t *n_vertices, int32_t **faces, int64_t *n_faces) {
  if (!bytes || !xyz || !n_vertices || !faces || !n_faces || len < 0) return fail(PTB_E_INVALID, "ply_parse_mesh: bad args");
  try {  // nothing may unwind through the C ABI into OCaml / Python / the CLI
    std::vector<float> v;
    std::vector<int32_t> f;
    int rc = parse((const unsigned char *)bytes, len, &v, &f);
    if (rc) return rc;
    return hand_over(v, f, xyz, n_vertices, faces, n_faces);
  } catch (const std::bad_alloc &) {
    return fail(PTB_E_NOMEM, "ply: out of memory");
  } catch (const std::exception &e) {
    return fail(PTB_E_INVALID, std::string("ply: ") + e.what());
  }
}

int ptb_ply_read_mesh(const char *path, float **xyz, int64_t *n_vertices, int32_t **faces, int64_t *n_faces) {
  if (!path) return fail(PTB_E_INVALID, "ply_read_mesh: null path");
  FILE *f = std::fopen(path, "rb");
  if (!f) return fail(PTB_E_INVALID, std::string("cannot open ") + path);
  std::fseek(f, 0, SEEK_END);
  long long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> buf;
  try {
    buf.resize((size_t)std::max<long long>(n, 0));
  } catch (const std::exception &) {
    std::fclose(f);
    return fail(PTB_E_NOMEM, "ply: out of memory");
  }
  size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  if ((long long)got != n) return fail(PTB_E_INVALID, std::string("short read on ") + path);
  return ptb_ply_parse_mesh(buf.data(), n, xyz, n_vertices, faces, n_faces);
}

void ptb_free(void *p) { std::free(p); }

}  // extern "C"
