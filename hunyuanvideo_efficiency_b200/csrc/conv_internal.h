// Internal (non-ABI) declarations shared by the tensor-core convolution translation units.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace hyvae {

template <typename T> struct TcFmt;
template <> struct TcFmt<__nv_bfloat16> { static constexpr int fmt = 1; };
template <> struct TcFmt<__half> { static constexpr int fmt = 0; };

// conv_halo.cu: stride-1 3x3x3 conv, Cout <= 128, full 2-D halo stage (see the header of that file)
struct HaloArgs {
  const float* bias;
  int B, To, Ho, Wo, Cin, Cout;
  int tiles_h, groups_w;   // 16-row tiles, groups of MT 8-column m-tiles
  int64_t total;           // B * To * tiles_h * groups_w
  int has_res;
  int sc_chunks, sc_cin;   // fused 1x1x1 shortcut conv: 64-channel chunks / channels of its input (0 = none)
  int round_like_ref;      // round conv + bias to the storage type before adding the residual
  double* gn_part;         // optional [B][gn_rows][gn_groups][2]
  int gn_groups, gn_cpg, gn_rows;
  int probe;
  int tfold;              // weights carry the 18 folded first-frame taps (tfold_* in tcgen05.cuh)
  int kwpack;             // THIN only: x holds the three kw taps along the channel axis, w is [9 = (kt, kh)][Cout][16]
};

void halo_geometry(int bn, int mt, bool pair, bool thin, int* twh, int* thh, int* taps_per_b, int* brows);
int launch_halo(int dtype, int bn, int mt, bool pair, bool thin, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                const CUtensorMap& tmR, const CUtensorMap& tmX, const CUtensorMap& tmW, const HaloArgs& a, cudaStream_t stream);

// rows per batch item of the GroupNorm partial buffer: one per (CTA, epilogue warp); the Winograd kernel has 8 epilogue warps
inline int gn_partial_rows() { return num_sms() * 8; }
// Layout of the caller's GroupNorm partial buffer (hyvae_gn_partials_doubles): [B][gn_partial_rows()][groups][2] per-warp rows,
// then [B][num_sms()][groups][2] per-CTA rows and one 8-byte ticket used by kernels that finish the statistics themselves.
inline int64_t gn_warp_rows_doubles(int B, int groups) { return (int64_t)B * gn_partial_rows() * groups * 2; }
inline int64_t gn_cta_rows_doubles(int B, int groups) { return (int64_t)B * num_sms() * groups * 2; }

// conv_wino.cu: stride-1 3x3x3 conv as Winograd F(2,3) along T over a plane volume (see the header of that file)
struct WinoArgs {
  const float* bias;
  int B, T, Ho, Wo, Cin, Cout;
  int tiles_h, groups_w;   // 16-row x 8-column m-tiles
  int gpf, upf;            // m-tiles per frame, CTA-pair units per frame (= ceil(gpf / 2))
  int npairs, ntu;         // output-frame pairs (T - 1) / 2; time units 1 + npairs (+ 1 for an even T)
  int n_tiles, total;      // 128-channel n-tiles; work items B * ntu * upf * n_tiles
  int has_res;
  int sc_chunks, sc_cin;   // fused 1x1x1 shortcut conv: 64-channel chunks / channels of its input (0 = none)
  double* gn_part;         // optional [B][gn_rows][gn_groups][2]
  int gn_groups, gn_cpg, gn_rows;
  double* gn_sums;         // optional [B][gn_groups][2]: the kernel finishes the statistics itself (last CTA sums in fixed order)
  double* gn_cta;          // [B][gn_cta_rows][gn_groups][2] per-CTA rows (inside the caller's partial buffer)
  unsigned int* gn_ticket; // arrival counter of the fold (zero between launches)
  int gn_cta_rows;
  int probe;               // measurement only (HYVAE_TC_PROBE): bit 0 = no weight loads once the ring is primed, bit 3 = no plane loads either
};
int launch_wino(int dtype, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                const CUtensorMap& tmX, const CUtensorMap& tmW, const WinoArgs& a, cudaStream_t stream);

// conv_stack.cu: stride-1 3x3x3 conv with Cout <= 8 (decoder conv_out), the nine (kh, kw) taps stacked along N.
// A box {64, 18, 18}; B box {64, 80, 1} over the packed weights viewed as [kt][72][Cin].
int launch_conv_stack(int dtype, const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloArgs& a, void* y, int64_t ysB, int64_t ysT,
                      int64_t ysH, int64_t ysW, int64_t yoff, cudaStream_t stream);

}  // namespace hyvae
