// Native Kaldi ark/scp reader and vector writer (host code, no CUDA): what the reference's Python
// kaldi_io does one byte at a time (pytorch/libs/support/kaldi_io.py `open_or_fd` :43-73, `read_key`
// :148-163, `_read_mat_binary` :495-525, `_read_compressed_mat` :527-569, `write_vec_flt` :367-399),
// for the Python-free extractor and as the fast reader in front of the GPU
// path (rank 1).  Formats:
//   ark entry : <key> ' ' + ('\0B' + binary payload | ascii " [ ... ]")
//   matrices  : 'FM '/'DM ' + \4 rows(i32) \4 cols(i32) + row-major data
//               'CM ' + {min f32, range f32, rows i32, cols i32} + cols x 4 u16 percentiles + cols x rows u8
//   vectors   : 'FV ' + \4 dim(i32) + data; text: " [ v v v ]\n"
//   rspecifier: ark:<file|-|cmd |>   scp:<file>   (scp lines: <key> <file[:offset] | cmd |>)
//   wspecifier: ark:<file|-|| cmd>   ark,scp:<ark>,<scp>   ark,t:<file>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/xvb200.h"

namespace xvb {
void set_error(const char* fmt, ...);
}
using xvb::set_error;

namespace {

struct Stream {
  FILE* f = nullptr;
  bool pipe = false;
  bool is_std = false;
  int close() {
    int rc = 0;
    if (f && !is_std) rc = pipe ? pclose(f) : fclose(f);
    else if (f) fflush(f);
    f = nullptr;
    return rc;
  }
};

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// "file", "file:123", "-", "cmd |", "| cmd"
bool open_stream(const std::string& spec_in, bool read, Stream* st, long* offset_out = nullptr) {
  std::string spec = trim(spec_in);
  long offset = -1;
  if (!spec.empty() && spec.back() != '|' && spec[0] != '|') {
    size_t c = spec.rfind(':');
    if (c != std::string::npos && c + 1 < spec.size() &&
        spec.find_first_not_of("0123456789", c + 1) == std::string::npos) {
      offset = atol(spec.c_str() + c + 1);
      spec = spec.substr(0, c);
    }
  }
  if (spec.empty()) { set_error("ark: empty file name"); return false; }
  if (read && spec.back() == '|') {
    st->f = popen(spec.substr(0, spec.size() - 1).c_str(), "r");
    st->pipe = true;
  } else if (!read && spec[0] == '|') {
    st->f = popen(spec.substr(1).c_str(), "w");
    st->pipe = true;
  } else if (spec == "-") {
    st->f = read ? stdin : stdout;
    st->is_std = true;
  } else {
    st->f = fopen(spec.c_str(), read ? "rb" : "wb");
  }
  if (!st->f) { set_error("ark: cannot open '%s'", spec.c_str()); return false; }
  if (offset >= 0 && fseek(st->f, offset, SEEK_SET) != 0) { set_error("ark: cannot seek '%s' to %ld", spec.c_str(), offset); return false; }
  if (offset_out) *offset_out = offset;
  return true;
}

bool read_exact(FILE* f, void* dst, size_t n) {
  if (fread(dst, 1, n, f) != n) { set_error("ark: unexpected end of stream (wanted %zu bytes)", n); return false; }
  return true;
}

bool read_dim(FILE* f, int32_t* v) {
  int c = fgetc(f);
  if (c != 4) { set_error("ark: expected int32 size marker, got %d", c); return false; }
  return read_exact(f, v, 4);
}

// 1 = key read, 0 = clean end of stream, -1 = error
int read_key(FILE* f, std::string* key) {
  key->clear();
  int c;
  while ((c = fgetc(f)) != EOF && c != ' ') key->push_back((char)c);
  *key = trim(*key);
  if (key->empty()) return c == EOF ? 0 : -1;
  if (key->find_first_of(" \t\n") != std::string::npos) { set_error("ark: malformed key"); return -1; }
  return 1;
}

// One matrix (binary FM/DM/CM or ascii) at the current position -> fp32 row-major.
bool read_matrix(FILE* f, std::vector<float>* out, int* rows, int* cols, std::vector<uint8_t>* scratch, bool* was_double = nullptr) {
  if (was_double) *was_double = false;
  char flag[2];
  if (!read_exact(f, flag, 2)) return false;
  if (flag[0] == '\0' && flag[1] == 'B') {
    char tok[3];
    if (!read_exact(f, tok, 3)) return false;
    if (tok[0] == 'C' && tok[1] == 'M' && tok[2] == ' ') {
      struct { float gmin, grange; int32_t rows, cols; } h;
      if (!read_exact(f, &h, 16)) return false;
      if (h.rows < 0 || h.cols < 0) { set_error("ark: negative compressed matrix size"); return false; }
      std::vector<uint16_t> perc((size_t)h.cols * 4);
      if (!read_exact(f, perc.data(), perc.size() * 2)) return false;
      scratch->resize((size_t)h.cols * h.rows);
      if (!read_exact(f, scratch->data(), scratch->size())) return false;
      out->resize((size_t)h.rows * h.cols);
      for (int c = 0; c < h.cols; ++c) {
        // same fp32 steps as the reference: u16 -> float: gmin + grange * 1.52590218966964e-05 * value
        float p[4];
        for (int k = 0; k < 4; ++k) {
          volatile float t = (float)perc[(size_t)c * 4 + k] * h.grange;
          volatile float u = t * 1.52590218966964e-05f;
          p[k] = u + h.gmin;
        }
        const uint8_t* col = scratch->data() + (size_t)c * h.rows;
        for (int r = 0; r < h.rows; ++r) {
          const uint8_t d = col[r];
          volatile float slope, prod;
          float v;
          if (d <= 64) { slope = (p[1] - p[0]) / 64.0f; prod = slope * (float)d; v = p[0] + prod; }
          else if (d > 192) { slope = (p[3] - p[2]) / 63.0f; prod = slope * (float)(d - 192); v = p[2] + prod; }
          else { slope = (p[2] - p[1]) / 128.0f; prod = slope * (float)(d - 64); v = p[1] + prod; }
          (*out)[(size_t)r * h.cols + c] = v;
        }
      }
      *rows = h.rows; *cols = h.cols;
      return true;
    }
    if (tok[0] == 'C' && tok[1] == 'M' && (tok[2] == '2' || tok[2] == '3')) {
      // Kaldi's two header-only compressed formats (restated from Kaldi's CompressedMatrix::Write/Read; Kaldi is not
      // vendored by the reference, so this is unpinned), which the reference's kaldi_io does not read: 'CM2' = uint16, 'CM3' = uint8, row-major, value = min + range * q / (65535 | 255).
      // A space follows the 3-character token.
      struct { float gmin, grange; int32_t rows, cols; } h;
      if (fgetc(f) != ' ' || !read_exact(f, &h, 16)) { set_error("ark: truncated CM%c header", tok[2]); return false; }
      if (h.rows < 0 || h.cols < 0) { set_error("ark: negative compressed matrix size"); return false; }
      const size_t n = (size_t)h.rows * h.cols;
      out->resize(n);
      if (tok[2] == '2') {
        scratch->resize(n * 2);
        if (!read_exact(f, scratch->data(), n * 2)) return false;
        const uint16_t* q = reinterpret_cast<const uint16_t*>(scratch->data());
        const float inc = h.grange * (1.0f / 65535.0f);
        for (size_t i = 0; i < n; ++i) (*out)[i] = h.gmin + (float)q[i] * inc;
      } else {
        scratch->resize(n);
        if (!read_exact(f, scratch->data(), n)) return false;
        const float inc = h.grange * (1.0f / 255.0f);
        for (size_t i = 0; i < n; ++i) (*out)[i] = h.gmin + (float)(*scratch)[i] * inc;
      }
      *rows = h.rows; *cols = h.cols;
      return true;
    }
    if (tok[0] == 'C' && tok[1] == 'M') { set_error("ark: unknown compressed format 'CM%c'", tok[2]); return false; }
    const bool dbl = tok[0] == 'D';
    if (was_double) *was_double = dbl;
    if (!((tok[0] == 'F' || dbl) && tok[1] == 'M' && tok[2] == ' ')) { set_error("ark: unknown matrix header '%c%c%c'", tok[0], tok[1], tok[2]); return false; }
    int32_t r, c;
    if (!read_dim(f, &r) || !read_dim(f, &c)) return false;
    if (r < 0 || c < 0) { set_error("ark: negative matrix size"); return false; }
    out->resize((size_t)r * c);
    if (!dbl) {
      if (!read_exact(f, out->data(), out->size() * 4)) return false;
    } else {
      scratch->resize(out->size() * 8);
      if (!read_exact(f, scratch->data(), scratch->size())) return false;
      const double* d = reinterpret_cast<const double*>(scratch->data());
      for (size_t i = 0; i < out->size(); ++i) (*out)[i] = (float)d[i];
    }
    *rows = r; *cols = c;
    return true;
  }
  if (flag[0] == ' ' && flag[1] == '[') {   // ascii: rows separated by newlines, closed by ']'
    out->clear();
    int r = 0, c = -1, cur = 0;
    std::string tok;
    for (;;) {
      int ch = fgetc(f);
      if (ch == EOF) { set_error("ark: end of stream inside an ascii matrix"); return false; }
      const bool sep = ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r';
      if (!sep && ch != ']') { tok.push_back((char)ch); continue; }
      if (!tok.empty()) { out->push_back(strtof(tok.c_str(), nullptr)); tok.clear(); ++cur; }
      if (ch == '\n' || ch == ']') {
        if (cur > 0) {
          if (c >= 0 && cur != c) { set_error("ark: ragged ascii matrix"); return false; }
          c = cur; ++r; cur = 0;
        }
        if (ch == ']') break;
      }
    }
    int ch = fgetc(f);
    if (ch != '\n' && ch != EOF) ungetc(ch, f);
    *rows = r; *cols = c < 0 ? 0 : c;
    return true;
  }
  set_error("ark: neither binary nor ascii matrix start (%d, %d)", (int)flag[0], (int)flag[1]);
  return false;
}

}  // namespace

struct xvb_ark_reader {
  bool scp = false;
  Stream in;               // the ark stream, or the scp list
  std::string key;
  std::vector<float> data;
  std::vector<uint8_t> scratch;
};

struct xvb_ark_writer {
  Stream ark;
  FILE* scp = nullptr;
  std::string ark_path;
  bool text = false;
  long long pos = 0;       // bytes written to the ark so far (scp offsets)
};

// "ark,t:xxx" -> kind "ark", opts ",t", rest "xxx"; no prefix -> kind "ark"
static void split_spec(const char* spec, std::string* kind, std::string* opts, std::string* rest) {
  std::string s = trim(spec ? spec : "");
  *kind = "ark"; opts->clear(); *rest = s;
  if (s.compare(0, 3, "ark") == 0 || s.compare(0, 3, "scp") == 0) {
    size_t c = s.find(':');
    if (c != std::string::npos && s.find_first_of(" /|") > c) {
      *kind = s.substr(0, 3);
      *opts = s.substr(3, c - 3);
      *rest = s.substr(c + 1);
    }
  }
}

extern "C" int xvb_ark_reader_open(xvb_ark_reader_t** out, const char* rspecifier) {
  if (!out || !rspecifier) { set_error("xvb_ark_reader_open: null argument"); return XVB_EINVAL; }
  std::string kind, opts, rest;
  split_spec(rspecifier, &kind, &opts, &rest);
  xvb_ark_reader* r = new xvb_ark_reader();
  if (opts.find(",scp") != std::string::npos && kind == "ark") kind = "scp";   // "ark,scp:" is a wspecifier form
  r->scp = kind == "scp";
  if (!open_stream(rest, true, &r->in)) { delete r; return XVB_EINVAL; }
  *out = r;
  return XVB_OK;
}

extern "C" int xvb_ark_reader_next(xvb_ark_reader_t* r, const char** key, int* rows, int* cols, const float** data) {
  if (!r || !key || !rows || !cols || !data) { set_error("xvb_ark_reader_next: null argument"); return XVB_EINVAL; }
  bool dbl = false;
  if (!r->scp) {
    const int k = read_key(r->in.f, &r->key);
    if (k == 0) return 0;
    if (k < 0 || !read_matrix(r->in.f, &r->data, rows, cols, &r->scratch, &dbl)) return XVB_EINVAL;
  } else {
    char* line = nullptr;
    size_t cap = 0;
    std::string l;
    for (;;) {
      if (getline(&line, &cap, r->in.f) < 0) { free(line); return 0; }
      l = trim(line);
      if (!l.empty()) break;
    }
    free(line);
    const size_t sp = l.find_first_of(" \t");
    if (sp == std::string::npos) { set_error("scp: line without a file: '%s'", l.c_str()); return XVB_EINVAL; }
    r->key = l.substr(0, sp);
    Stream st;
    if (!open_stream(l.substr(sp + 1), true, &st)) return XVB_EINVAL;
    const bool ok = read_matrix(st.f, &r->data, rows, cols, &r->scratch, &dbl);
    st.close();
    if (!ok) return XVB_EINVAL;
  }
  *key = r->key.c_str();
  *data = r->data.data();
  return dbl ? 2 : 1;   // 2: stored as a double ('DM ') matrix, converted
}

extern "C" void xvb_ark_reader_close(xvb_ark_reader_t* r) {
  if (!r) return;
  r->in.close();
  delete r;
}

extern "C" int xvb_ark_writer_open(xvb_ark_writer_t** out, const char* wspecifier) {
  if (!out || !wspecifier) { set_error("xvb_ark_writer_open: null argument"); return XVB_EINVAL; }
  std::string kind, opts, rest;
  split_spec(wspecifier, &kind, &opts, &rest);
  if (kind != "ark") { set_error("xvb_ark_writer_open: wspecifier must start with ark[,scp][,t]: ('%s')", wspecifier); return XVB_EINVAL; }
  xvb_ark_writer* w = new xvb_ark_writer();
  w->text = opts.find(",t") != std::string::npos;
  std::string ark = rest, scp;
  if (opts.find(",scp") != std::string::npos) {
    const size_t c = rest.find(',');
    if (c == std::string::npos) { set_error("xvb_ark_writer_open: ark,scp: needs <ark>,<scp>"); delete w; return XVB_EINVAL; }
    ark = trim(rest.substr(0, c));
    scp = trim(rest.substr(c + 1));
  }
  if (!open_stream(ark, false, &w->ark)) { delete w; return XVB_EINVAL; }
  w->ark_path = trim(ark);
  if (!scp.empty()) {
    if (w->ark.pipe || w->ark.is_std) { set_error("xvb_ark_writer_open: scp offsets need a regular ark file"); w->ark.close(); delete w; return XVB_EINVAL; }
    w->scp = fopen(scp.c_str(), "w");
    if (!w->scp) { set_error("xvb_ark_writer_open: cannot open '%s'", scp.c_str()); w->ark.close(); delete w; return XVB_EINVAL; }
  }
  *out = w;
  return XVB_OK;
}

extern "C" int xvb_ark_writer_put_vector(xvb_ark_writer_t* w, const char* key, const float* v, int dim) {
  if (!w || !key || (!v && dim > 0) || dim < 0) { set_error("xvb_ark_writer_put_vector: bad argument"); return XVB_EINVAL; }
  if (!key[0] || strpbrk(key, " \t\n")) { set_error("xvb_ark_writer_put_vector: key must be one non-empty token"); return XVB_EINVAL; }
  FILE* f = w->ark.f;
  const size_t klen = strlen(key);
  bool ok = fwrite(key, 1, klen, f) == klen && fputc(' ', f) != EOF;
  w->pos += (long long)klen + 1;
  if (w->scp) fprintf(w->scp, "%s %s:%lld\n", key, w->ark_path.c_str(), w->pos);
  if (w->text) {
    std::string s = " [ ";
    char buf[48];
    for (int i = 0; i < dim; ++i) { snprintf(buf, sizeof buf, "%.9g ", (double)v[i]); s += buf; }
    s += "]\n";
    ok = ok && fwrite(s.data(), 1, s.size(), f) == s.size();
    w->pos += (long long)s.size();
  } else {
    const char head[6] = {'\0', 'B', 'F', 'V', ' ', '\4'};
    const int32_t d = dim;
    ok = ok && fwrite(head, 1, 6, f) == 6 && fwrite(&d, 4, 1, f) == 1 &&
         (dim == 0 || fwrite(v, 4, (size_t)dim, f) == (size_t)dim);
    w->pos += 10 + 4LL * dim;
  }
  if (!ok) { set_error("xvb_ark_writer_put_vector: write failed for key '%s'", key); return XVB_EINVAL; }
  return XVB_OK;
}

extern "C" int xvb_ark_writer_close(xvb_ark_writer_t* w) {
  if (!w) return XVB_OK;
  int rc = w->ark.close();
  if (w->scp && fclose(w->scp) != 0) rc = -1;
  delete w;
  if (rc != 0) { set_error("xvb_ark_writer_close: stream or pipe command failed (%d)", rc); return XVB_EINVAL; }
  return XVB_OK;
}
