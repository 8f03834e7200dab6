// conv_api.cu -- argument validation and algorithm choice for the convolution entry points.
#include "conv_common.cuh"

using namespace mvd;

static int check_geom(const mvd_conv3d_args* a, const char* who) {
  MVD_REQUIRE(a != nullptr, "%s: null args", who);
  MVD_REQUIRE(a->B > 0 && a->Di > 0 && a->Hi > 0 && a->Wi > 0 && a->Cin > 0 && a->Do > 0 && a->Ho > 0 && a->Wo > 0 &&
                  a->Cout > 0, "%s: non-positive dimension", who);
  MVD_REQUIRE(a->kd > 0 && a->kh > 0 && a->kw > 0 && a->sd > 0 && a->sh > 0 && a->sw > 0 && a->pd >= 0 && a->ph >= 0 &&
                  a->pw >= 0, "%s: bad kernel/stride/padding", who);
  MVD_REQUIRE(a->Do == (a->Di + 2 * a->pd - a->kd) / a->sd + 1 && a->Ho == (a->Hi + 2 * a->ph - a->kh) / a->sh + 1 &&
                  a->Wo == (a->Wi + 2 * a->pw - a->kw) / a->sw + 1,
              "%s: output extent does not match floor((in + 2p - k)/s) + 1", who);
  MVD_REQUIRE(a->x && a->y && a->ldx >= a->Cin && a->ldy >= a->Cout, "%s: bad activation pointers / pitches", who);
  MVD_REQUIRE(a->algo >= 0 && a->algo <= 2, "%s: algo must be 0, 1 or 2", who);
  return MVD_OK;
}

extern "C" {

int mvd_set_deterministic(int on) {
  set_wgrad_deterministic(on);
  return MVD_OK;
}
int mvd_get_deterministic(void) { return wgrad_deterministic() ? 1 : 0; }

size_t mvd_conv3d_workspace_bytes(const mvd_conv3d_args* a, int pass) {
  if (!a) return 0;
  if (pass == 2 && a->algo != 1 && tc_wgrad_supported(a)) return tc_wgrad_workspace_bytes(a);
  if ((pass == 0 || pass == 1) && a->algo != 1 && a->x && a->y && a->w) return tc_splitk_workspace_bytes(a, pass);
  return 0;
}

int mvd_conv3d_fprop(const mvd_conv3d_args* a, mvd_stream_t stream) {
  int rc = check_geom(a, "conv3d_fprop");
  if (rc) return rc;
  MVD_REQUIRE(a->w, "conv3d_fprop: null weights");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = tc_fprop_supported(a);
  if (a->algo == 2 && !tc) { set_error("conv3d_fprop: shape not covered by the tcgen05 kernel"); return MVD_ERR_UNSUPPORTED; }
  const bool use_tc = (a->algo != 1 && tc);
  if (a->algo == 0 && !tc) count_fallback();
  // InstanceNorm statistics: fused into the tcgen05 epilogue where that epilogue has slack (narrow layers, which are
  // also the ones with the most voxels); a separate streaming pass otherwise (measured: for N >= 128 the extra
  // shuffle-reduction per 32-column group makes the epilogue the bottleneck and costs more than the pass it saves)
  const bool fuse_stats = use_tc && a->stats && (a->Cout == 32 || a->Cout == 64);
  mvd_conv3d_args b = *a;
  if (!fuse_stats) b.stats = nullptr;
  rc = use_tc ? tc_fprop(&b, st) : generic_fprop(&b, st);
  if (rc) return rc;
  if (a->stats && !fuse_stats)
    return mvd_inorm_stats(a->y, a->ldy, a->B, (long long)a->Do * a->Ho * a->Wo, a->Cout, a->stats, stream);
  return MVD_OK;
}

int mvd_conv3d_dgrad(const mvd_conv3d_args* a, mvd_stream_t stream) {
  int rc = check_geom(a, "conv3d_dgrad");
  if (rc) return rc;
  MVD_REQUIRE(a->w, "conv3d_dgrad: null weights");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = tc_dgrad_supported(a);
  if (a->algo == 2 && !tc) { set_error("conv3d_dgrad: shape not covered by the tcgen05 kernel"); return MVD_ERR_UNSUPPORTED; }
  if (a->algo == 0 && !tc) count_fallback();
  // optional sums of the produced gradient (stats = [B][Cin][2]: sum, sum of squares): fused into the halo kernel's
  // epilogue for 3x3x3 / stride-1 layers with 32 or 64 input channels, a streaming pass otherwise
  const bool k3s1p1 = a->kd == 3 && a->kh == 3 && a->kw == 3 && a->sd == 1 && a->sh == 1 && a->sw == 1 && a->pd == 1 &&
                      a->ph == 1 && a->pw == 1;
  const bool fuse = a->stats && tc && a->algo != 1 && k3s1p1 && !a->accumulate && (a->Cin == 32 || a->Cin == 64) &&
                    tc_splitk_workspace_bytes(a, 1) == 0;
  // optional backward statistics of the InstanceNorm + LeakyReLU in front of x: halo-kernel epilogue, else a pass over x
  const bool fuse_nb = a->norm_bwd && a->algo != 1 && tc && tc_dgrad_fuses_norm_bwd(a);
  mvd_conv3d_args b = *a;
  if (!fuse) b.stats = nullptr;
  if (!fuse_nb) b.norm_bwd = nullptr;
  rc = (a->algo != 1 && tc) ? tc_dgrad(&b, st) : generic_dgrad(&b, st);
  if (rc) return rc;
  if (a->stats && !fuse) {
    rc = mvd_inorm_stats(a->x, a->ldx, a->B, (long long)a->Di * a->Hi * a->Wi, a->Cin, a->stats, stream);
    if (rc) return rc;
  }
  if (a->norm_bwd && !fuse_nb) {
    const mvd_norm_bwd_stats_args* nb = a->norm_bwd;
    MVD_REQUIRE(nb->y && nb->stats && nb->bstats, "conv3d_dgrad: incomplete norm_bwd description");
    return mvd_inorm_lrelu_bwd_stats(a->x, a->ldx, nb->y, nb->ldy, nb->stats, nb->gamma, nb->beta, a->B,
                                     (long long)a->Di * a->Hi * a->Wi, a->Cin, nb->eps, nb->slope, nb->bstats, stream);
  }
  return MVD_OK;
}

int mvd_conv3d_dgrad_fuses_norm_bwd(const mvd_conv3d_args* a) {
  if (!a || !a->norm_bwd || !a->x || !a->y || !a->w || a->algo == 1) return 0;
  return tc_dgrad_fuses_norm_bwd(a) ? 1 : 0;
}

int mvd_conv3d_wgrad(const mvd_conv3d_args* a, mvd_stream_t stream) {
  int rc = check_geom(a, "conv3d_wgrad");
  if (rc) return rc;
  MVD_REQUIRE(a->dw, "conv3d_wgrad: null dw");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = tc_wgrad_supported(a);
  if (a->algo == 2 && !tc) { set_error("conv3d_wgrad: shape not covered by the tcgen05 kernel"); return MVD_ERR_UNSUPPORTED; }
  if (a->algo == 0 && !tc) count_fallback();
  return (a->algo != 1 && tc) ? tc_wgrad(a, st) : generic_wgrad(a, st);
}

}  // extern "C"
