// .xvbm model files: the layer list of a TDNN-family extractor exactly as the reference stores it in
// its state_dict (conv weights (Cout, Cin, tot_context) incl. masked taps, eval BatchNorm folded to
// scale/shift), written by asv_subtools_b200.ops.Extractor.save() and loaded here without Python --
// the role of torch::jit::load in the reference's runtime (runtime/extractor/torch_asv_model.cc:8-17).
//
//   "XVBM0001" | i32 feat_dim | f32 pooling_eps | i32 n_frame | i32 n_segment
//   frame layer  : i32 Cout, Cin, ntaps, tot_context, flags, has_bias, has_bn | i32 ctx[ntaps]
//                  | f32 w[Cout*Cin*tot_context] | f32 bias[Cout]? | f32 scale[Cout], shift[Cout]?
//   segment layer: same header with ntaps = tot_context = 1 and ctx = {0}
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/xvb200.h"

namespace xvb {
void set_error(const char* fmt, ...);
}
using xvb::set_error;

namespace {
bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }
}

extern "C" int xvb_extractor_load(xvb_extractor_t** out, const char* path) {
  if (!out || !path) { set_error("xvb_extractor_load: null argument"); return XVB_EINVAL; }
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("xvb_extractor_load: cannot open '%s'", path); return XVB_EINVAL; }
  xvb_extractor_t* h = nullptr;
  int rc = XVB_EINVAL;
  char magic[8];
  int32_t feat_dim = 0, n_frame = 0, n_seg = 0;
  float eps = 0.f;
  do {
    if (!rd(f, magic, 8) || memcmp(magic, "XVBM0001", 8) != 0) { set_error("xvb_extractor_load: '%s' is not an XVBM0001 file", path); break; }
    if (!rd(f, &feat_dim, 4) || !rd(f, &eps, 4) || !rd(f, &n_frame, 4) || !rd(f, &n_seg, 4) || feat_dim <= 0 ||
        n_frame <= 0 || n_seg <= 0 || n_frame > 64 || n_seg > 64) { set_error("xvb_extractor_load: bad header in '%s'", path); break; }
    if ((rc = xvb_extractor_create(&h, feat_dim)) != XVB_OK) break;
    std::vector<float> w, bias, scale, shift;
    bool ok = true;
    for (int i = 0; ok && i < n_frame + n_seg; ++i) {
      int32_t hd[7];
      int32_t ctx[XVB_MAX_TAPS];
      ok = rd(f, hd, sizeof hd);
      const int Cout = hd[0], Cin = hd[1], ntaps = hd[2], tot = hd[3], flags = hd[4], has_bias = hd[5], has_bn = hd[6];
      ok = ok && Cout > 0 && Cin > 0 && ntaps >= 1 && ntaps <= XVB_MAX_TAPS && tot >= ntaps && tot < 4096 && rd(f, ctx, 4 * (size_t)ntaps);
      if (!ok) { set_error("xvb_extractor_load: bad layer %d header in '%s'", i, path); rc = XVB_EINVAL; break; }
      w.resize((size_t)Cout * Cin * tot);
      ok = rd(f, w.data(), w.size() * 4);
      if (ok && has_bias) { bias.resize(Cout); ok = rd(f, bias.data(), 4 * (size_t)Cout); }
      if (ok && has_bn) { scale.resize(Cout); shift.resize(Cout); ok = rd(f, scale.data(), 4 * (size_t)Cout) && rd(f, shift.data(), 4 * (size_t)Cout); }
      if (!ok) { set_error("xvb_extractor_load: '%s' is truncated in layer %d", path, i); rc = XVB_EINVAL; break; }
      if (i < n_frame)
        rc = xvb_extractor_add_frame_layer(h, Cout, ctx, ntaps, w.data(), has_bias ? bias.data() : nullptr,
                                           has_bn ? scale.data() : nullptr, has_bn ? shift.data() : nullptr, flags);
      else
        rc = xvb_extractor_add_segment_layer(h, Cout, w.data(), has_bias ? bias.data() : nullptr,
                                             has_bn ? scale.data() : nullptr, has_bn ? shift.data() : nullptr, flags);
      ok = rc == XVB_OK;
    }
    if (!ok) break;
    rc = xvb_extractor_finalize(h, eps);
  } while (0);
  fclose(f);
  if (rc != XVB_OK) { if (h) xvb_extractor_destroy(h); return rc; }
  *out = h;
  return XVB_OK;
}

extern "C" int xvb_extractor_feat_dim(const char* path) {
  FILE* f = path ? fopen(path, "rb") : nullptr;
  if (!f) { set_error("xvb_extractor_feat_dim: cannot open '%s'", path ? path : "(null)"); return XVB_EINVAL; }
  char magic[8];
  int32_t d = 0;
  const bool ok = rd(f, magic, 8) && memcmp(magic, "XVBM0001", 8) == 0 && rd(f, &d, 4) && d > 0;
  fclose(f);
  if (!ok) { set_error("xvb_extractor_feat_dim: '%s' is not an XVBM0001 file", path); return XVB_EINVAL; }
  return d;
}
