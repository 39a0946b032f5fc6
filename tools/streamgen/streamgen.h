// streamgen.h — options of the test-only bitstream writers.
#pragma once
#include <stdint.h>
#include <vector>

namespace sg {

struct H264Opts {
	int mbw = 8, mbh = 6, frames = 8, seed = 1;
	int profile = 66;          // Baseline
	int gop = 1000;            // IDR period in frames
	int refs = 1;              // num_ref_frames (and default active count)
	int qp = 28;
	int chroma_qp_offset = 0;
	int vary_qp = 1;           // slice_qp_delta / mb_qp_delta within +-4 of qp
	int deblock = 1;           // 0: filter off (idc 1); 1: on (idc 0); 2: off in odd slices
	int alpha_div2 = 0, beta_div2 = 0;   // keep beta >= alpha (hazard 5)
	int slices = 1;
	int intra_pct = 10, skip_pct = 15, ipcm_pct = 1, coded_pct = 25;
	int max_level = 3, big_levels = 0;
	int dense_pct = 0;         // share of coded 4x4 blocks that are nearly full (13-16 coefficients)
	int amp_budget = 200;      // bound on the L1 amplitude of a block's residual (hazards 1, 2)
	int mv_step = 6, mv_range = 96, mv_out = 24;   // quarter-pel step / range, pixels allowed outside
	int crop_bottom = 0, crop_right = 0;
	int constrained_intra = 0;
	// Main-profile features (CAVLC): B pictures between anchors, direct modes, weighted prediction
	int bframes = 0;           // non-reference B pictures between consecutive P/I pictures (POC type 0)
	int direct_spatial = 1;    // direct_spatial_mv_pred_flag
	int direct_pct = 20;       // share of B macroblocks / sub-macroblocks coded as direct, and B_Skip share = skip_pct
	int weighted_pred = 0;     // explicit weights in P slices
	int weighted_bipred = 0;   // 0 default, 1 explicit, 2 implicit
	int direct_8x8_inference = 1;
	// entropy coder / High-profile tools
	int cabac = 0;             // entropy_coding_mode_flag
	int cabac_init_idc = -1;   // -1: drawn per slice
	int transform8x8 = 0;      // transform_8x8_mode_flag (needs cabac = 1 and profile = 100): Intra8x8 + 8x8 residual
};

struct Mpeg2Opts {
	int mbw = 8, mbh = 6, frames = 9, seed = 1;
	int gop = 12, bframes = 2;
	int qscale = 6;
	int intra_pct = 8, skip_pct = 10, coded_pct = 40;
	int field = 0;             // 1: frame pictures with field prediction / field DCT allowed
	int intra_vlc = 0, alt_scan = 0, q_scale_type = 0, dc_precision = 0;
	int mv_range = 24;         // half-pel
	int width = 0, height = 0; // 0: mbw*16 x mbh*16
};

std::vector<uint8_t> write_h264(const H264Opts &o);
std::vector<uint8_t> write_mpeg2(const Mpeg2Opts &o);

}  // namespace sg
