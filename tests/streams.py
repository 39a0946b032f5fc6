"""Named test streams: writer arguments (tools/streamgen) for each parity case."""
import os
import subprocess

import util

STREAMGEN = os.path.join(util.ROOT, "tools", "_bin", "streamgen")

# name -> (codec, args).  Small enough for the CPU oracle / reference to finish in seconds.
H264_CASES = {
    "h264_basic": "mbw=8 mbh=6 frames=6 seed=1",
    "h264_nodeblock": "mbw=8 mbh=6 frames=6 seed=2 deblock=0",
    "h264_refs3": "mbw=11 mbh=9 frames=12 seed=3 refs=3",
    "h264_gop7": "mbw=11 mbh=9 frames=20 seed=4 refs=2 gop=7",
    "h264_slices3": "mbw=10 mbh=7 frames=8 seed=5 slices=3",
    "h264_slices_dbk_mix": "mbw=10 mbh=7 frames=8 seed=6 slices=4 deblock=2",
    "h264_intra_ipcm": "mbw=6 mbh=5 frames=10 seed=7 intra_pct=50 ipcm_pct=10",
    "h264_skippy": "mbw=6 mbh=5 frames=10 seed=8 skip_pct=60",
    "h264_qp20_dense": "mbw=9 mbh=5 frames=8 seed=9 qp=20 coded_pct=60",
    "h264_qp36_ab": "mbw=9 mbh=5 frames=8 seed=10 qp=36 alpha_div2=-2 beta_div2=1",
    "h264_qp40_cqp": "mbw=9 mbh=5 frames=8 seed=11 qp=40 alpha_div2=3 beta_div2=4 chroma_qp_offset=4",
    "h264_umv": "mbw=9 mbh=5 frames=8 seed=12 qp=30 chroma_qp_offset=-6 mv_range=300 mv_out=60",
    "h264_big": "mbw=20 mbh=12 frames=24 seed=13 refs=4 gop=20 slices=2",
    "h264_crop": "mbw=5 mbh=4 frames=6 seed=14 crop_bottom=8 crop_right=6",
    "h264_long": "mbw=8 mbh=8 frames=40 seed=15 refs=2",
    "h264_biglevels": "mbw=8 mbh=6 frames=6 seed=16 qp=12 big_levels=1 coded_pct=50",
    "h264_constrained": "mbw=8 mbh=6 frames=8 seed=17 constrained_intra=1 intra_pct=30",
    "h264_2x2": "mbw=2 mbh=2 frames=6 seed=21",
    # nearly full 4x4 blocks (13-16 coefficients): the long coeff_token codes and total_zeros tables 13-15
    # (round 2: total_zeros of 15 coefficients was inverted in writer and parser alike; the reference stopped)
    "h264_full_blocks": "mbw=12 mbh=8 frames=6 seed=5 dense_pct=30",
    "h264_full_blocks_intra": "mbw=12 mbh=8 frames=6 seed=6 dense_pct=60 coded_pct=60 intra_pct=30",
    "h264_full_blocks_b": "mbw=12 mbh=8 frames=8 seed=7 dense_pct=40 bframes=2 profile=77 refs=2",
    # Main profile, CAVLC: B pictures, direct modes, weighted prediction
    "h264_b_bipred": "mbw=8 mbh=6 frames=7 seed=31 refs=2 bframes=2 profile=77 skip_pct=0 direct_pct=0",
    "h264_b_implicit": "mbw=8 mbh=6 frames=7 seed=32 refs=2 bframes=2 profile=77 skip_pct=0 direct_pct=0 weighted_bipred=2",
    "h264_b_explicit": "mbw=8 mbh=6 frames=7 seed=33 refs=2 bframes=2 profile=77 skip_pct=0 direct_pct=0 weighted_bipred=1 weighted_pred=1",
    "h264_b_spatial": "mbw=8 mbh=6 frames=7 seed=34 refs=2 bframes=2 profile=77",
    "h264_b_temporal": "mbw=8 mbh=6 frames=7 seed=35 refs=2 bframes=2 profile=77 direct_spatial=0",
    "h264_b_temporal_implicit": "mbw=10 mbh=8 frames=13 seed=36 refs=3 bframes=2 profile=77 direct_spatial=0 weighted_bipred=2",
    "h264_b3_explicit_slices": "mbw=10 mbh=8 frames=13 seed=37 refs=3 bframes=3 profile=77 weighted_bipred=1 weighted_pred=1 slices=2",
    "h264_b_direct4x4_spatial": "mbw=10 mbh=8 frames=10 seed=38 refs=2 bframes=1 profile=77 direct_8x8_inference=0",
    "h264_b_direct4x4_temporal": "mbw=10 mbh=8 frames=10 seed=39 refs=2 bframes=2 profile=77 direct_8x8_inference=0 direct_spatial=0",
    "h264_b_mixed": "mbw=12 mbh=9 frames=16 seed=40 refs=4 bframes=2 profile=77 intra_pct=20 direct_pct=40 skip_pct=30",
    "h264_p_weighted": "mbw=9 mbh=6 frames=8 seed=42 refs=3 weighted_pred=1 profile=77",
    "h264_2x3_slices": "mbw=2 mbh=3 frames=6 seed=25 slices=2",
    # Main profile, CABAC
    "h264_cabac_ip": "mbw=8 mbh=6 frames=8 seed=51 cabac=1 profile=77 refs=3",
    "h264_cabac_intra_ipcm": "mbw=8 mbh=6 frames=10 seed=52 cabac=1 profile=77 intra_pct=50 ipcm_pct=10",
    "h264_cabac_slices": "mbw=10 mbh=7 frames=8 seed=53 cabac=1 profile=77 slices=3 refs=2 deblock=2",
    "h264_cabac_dense": "mbw=9 mbh=5 frames=8 seed=54 cabac=1 profile=77 qp=20 coded_pct=60",
    "h264_cabac_biglevels": "mbw=8 mbh=6 frames=6 seed=55 cabac=1 profile=77 qp=12 big_levels=1 coded_pct=50 cabac_init_idc=2",
    "h264_cabac_umv": "mbw=9 mbh=5 frames=8 seed=56 cabac=1 profile=77 qp=30 chroma_qp_offset=-6 mv_range=300 mv_out=60",
    "h264_cabac_b_spatial": "mbw=8 mbh=6 frames=7 seed=57 cabac=1 profile=77 refs=2 bframes=2",
    "h264_cabac_b_temporal_implicit": "mbw=10 mbh=8 frames=13 seed=58 cabac=1 profile=77 refs=3 bframes=2 direct_spatial=0 weighted_bipred=2",
    "h264_cabac_b3_explicit_slices": "mbw=10 mbh=8 frames=13 seed=59 cabac=1 profile=77 refs=3 bframes=3 weighted_bipred=1 weighted_pred=1 slices=2",
    "h264_cabac_b_direct4x4": "mbw=10 mbh=8 frames=10 seed=60 cabac=1 profile=77 refs=2 bframes=1 direct_8x8_inference=0",
    "h264_cabac_b_mixed": "mbw=12 mbh=9 frames=16 seed=61 cabac=1 profile=77 refs=4 bframes=2 intra_pct=20 direct_pct=40 skip_pct=30",
    "h264_cabac_full_blocks": "mbw=12 mbh=8 frames=6 seed=8 dense_pct=40 cabac=1 profile=77",
    "h264_cabac_skippy": "mbw=6 mbh=5 frames=10 seed=62 cabac=1 profile=77 skip_pct=60 cabac_init_idc=1",
    # High profile: CABAC + 8x8 transform + Intra8x8
    "h264_high_i": "mbw=8 mbh=6 frames=2 seed=71 cabac=1 profile=100 transform8x8=1 gop=1",
    "h264_high_ip": "mbw=8 mbh=6 frames=8 seed=72 cabac=1 profile=100 transform8x8=1 refs=2",
    "h264_high_b": "mbw=10 mbh=8 frames=10 seed=73 cabac=1 profile=100 transform8x8=1 refs=2 bframes=2",
    "h264_high_b_temporal_w": "mbw=10 mbh=8 frames=13 seed=74 cabac=1 profile=100 transform8x8=1 refs=3 bframes=2 direct_spatial=0 weighted_bipred=2 weighted_pred=1",
    "h264_high_slices_dbk": "mbw=10 mbh=7 frames=8 seed=75 cabac=1 profile=100 transform8x8=1 refs=2 slices=4 deblock=2",
    "h264_high_intra_ipcm": "mbw=6 mbh=5 frames=10 seed=76 cabac=1 profile=100 transform8x8=1 intra_pct=50 ipcm_pct=10",
    "h264_high_qp40": "mbw=9 mbh=5 frames=8 seed=77 cabac=1 profile=100 transform8x8=1 qp=40 alpha_div2=3 beta_div2=4 chroma_qp_offset=4",
    "h264_high_qp8_trunc": "mbw=8 mbh=6 frames=6 seed=78 cabac=1 profile=100 transform8x8=1 qp=8 coded_pct=50 vary_qp=0",
    "h264_high_big": "mbw=20 mbh=12 frames=24 seed=79 cabac=1 profile=100 transform8x8=1 refs=4 gop=20 slices=2 bframes=1",
    "h264_high_umv_dense": "mbw=9 mbh=5 frames=8 seed=80 cabac=1 profile=100 transform8x8=1 qp=22 coded_pct=60 mv_range=300 mv_out=60",
    "h264_high_full_blocks": "mbw=12 mbh=8 frames=6 seed=9 dense_pct=40 cabac=1 profile=100 transform8x8=1",
    "h264_high_2x2": "mbw=2 mbh=2 frames=6 seed=81 cabac=1 profile=100 transform8x8=1",
}
# Pictures one macroblock wide or high: the compiled reference disagrees with ffmpeg there (its
# own bug), our decode agrees with ffmpeg.  Used for CUDA-vs-oracle (and ffmpeg luma) only.
H264_REFERENCE_DEVIATES = {
    "h264_1mb": "mbw=1 mbh=1 frames=5 seed=18",
    "h264_row": "mbw=7 mbh=1 frames=5 seed=19",
    "h264_col": "mbw=1 mbh=7 frames=5 seed=20 slices=2",
}
# constrained_intra_pred with CABAC: the reference takes the ctxIdxInc of coded_block_flag from the
# constrained-intra-masked availability (h264.cpp:4093-4118 pass avail_intra to ResidualBlock; the
# standard does so only under data partitioning), loses the arithmetic decoder's sync and stops.
# Nothing to be identical to: CUDA vs oracle only.
H264_ORACLE_ONLY = {
    "h264_cabac_constrained": "mbw=8 mbh=6 frames=8 seed=63 cabac=1 profile=77 constrained_intra=1 intra_pct=30",
    "h264_high_constrained": "mbw=8 mbh=6 frames=8 seed=82 cabac=1 profile=100 transform8x8=1 constrained_intra=1 intra_pct=30",
}
# committed under tests/golden/ together with the reference's MD5 lines
H264_GOLDEN = ["h264_basic", "h264_full_blocks", "h264_refs3", "h264_slices_dbk_mix", "h264_intra_ipcm", "h264_qp40_cqp", "h264_umv", "h264_constrained",
               "h264_b_spatial", "h264_b_temporal_implicit", "h264_b3_explicit_slices", "h264_b_direct4x4_temporal", "h264_p_weighted",
               "h264_cabac_ip", "h264_cabac_intra_ipcm", "h264_cabac_b3_explicit_slices", "h264_cabac_b_mixed",
               "h264_high_ip", "h264_high_b_temporal_w", "h264_high_slices_dbk", "h264_high_intra_ipcm", "h264_high_qp8_trunc"]

MPEG2_CASES = {
    "mpeg2_ipb": "mbw=8 mbh=6 frames=7 seed=1",
    "mpeg2_ip": "mbw=8 mbh=6 frames=10 seed=2 bframes=0",
    "mpeg2_field": "mbw=10 mbh=7 frames=13 seed=3 field=1",
    "mpeg2_b15_altscan": "mbw=6 mbh=5 frames=9 seed=4 intra_vlc=1 alt_scan=1",
    "mpeg2_nonlinear_q": "mbw=6 mbh=5 frames=9 seed=5 q_scale_type=1 dc_precision=2 qscale=10",
    "mpeg2_720x480": "mbw=45 mbh=30 frames=7 seed=6",
    "mpeg2_b3_field_intra": "mbw=9 mbh=6 frames=12 seed=7 bframes=3 gop=6 field=1 intra_pct=30 skip_pct=30",
    "mpeg2_crop": "mbw=8 mbh=6 frames=7 seed=8 width=120 height=90",
    "mpeg2_2x2": "mbw=2 mbh=2 frames=7 seed=9",
}
MPEG2_GOLDEN = ["mpeg2_ipb", "mpeg2_field", "mpeg2_b15_altscan", "mpeg2_b3_field_intra", "mpeg2_crop"]


def generate(codec, args):
    """Run the test-only writer; returns the elementary stream bytes."""
    out = subprocess.run([STREAMGEN, codec, *args.split()], stdout=subprocess.PIPE, check=True)
    return out.stdout
