// streamgen — test-only elementary-stream generator.
//   streamgen h264  -o out.264 [key=value ...]
//   streamgen mpeg2 -o out.m2v [key=value ...]
// Keys are the field names of sg::H264Opts / sg::Mpeg2Opts (tools/streamgen/streamgen.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include "streamgen.h"

int main(int argc, char **argv) {
	if (argc < 2) { fprintf(stderr, "usage: streamgen h264|mpeg2 -o file key=value...\n"); return 2; }
	std::string codec = argv[1], out;
	std::map<std::string, int> kv;
	for (int i = 2; i < argc; ++i) {
		if (!strcmp(argv[i], "-o") && i + 1 < argc) { out = argv[++i]; continue; }
		const char *eq = strchr(argv[i], '=');
		if (!eq) { fprintf(stderr, "bad argument %s\n", argv[i]); return 2; }
		kv[std::string(argv[i], eq - argv[i])] = atoi(eq + 1);
	}
	std::vector<uint8_t> data;
#define OPT(name) if (kv.count(#name)) { o.name = kv[#name]; kv.erase(#name); }
	if (codec == "h264") {
		sg::H264Opts o;
		OPT(mbw) OPT(mbh) OPT(frames) OPT(seed) OPT(profile) OPT(gop) OPT(refs) OPT(qp) OPT(chroma_qp_offset)
		OPT(vary_qp) OPT(deblock) OPT(alpha_div2) OPT(beta_div2) OPT(slices) OPT(intra_pct) OPT(skip_pct)
		OPT(ipcm_pct) OPT(coded_pct) OPT(max_level) OPT(big_levels) OPT(dense_pct) OPT(amp_budget) OPT(mv_step) OPT(mv_range)
		OPT(mv_out) OPT(crop_bottom) OPT(crop_right) OPT(constrained_intra)
		OPT(bframes) OPT(direct_spatial) OPT(direct_pct) OPT(weighted_pred) OPT(weighted_bipred) OPT(direct_8x8_inference)
		OPT(cabac) OPT(cabac_init_idc) OPT(transform8x8)
		data = sg::write_h264(o);
	} else if (codec == "mpeg2") {
		sg::Mpeg2Opts o;
		OPT(mbw) OPT(mbh) OPT(frames) OPT(seed) OPT(gop) OPT(bframes) OPT(qscale) OPT(intra_pct) OPT(skip_pct)
		OPT(coded_pct) OPT(field) OPT(intra_vlc) OPT(alt_scan) OPT(q_scale_type) OPT(dc_precision) OPT(mv_range)
		OPT(width) OPT(height)
		data = sg::write_mpeg2(o);
	} else {
		fprintf(stderr, "unknown codec %s\n", codec.c_str());
		return 2;
	}
	for (auto &p : kv) { fprintf(stderr, "unknown option %s\n", p.first.c_str()); return 2; }
	FILE *f = out.empty() ? stdout : fopen(out.c_str(), "wb");
	if (!f) { perror("open"); return 1; }
	fwrite(data.data(), 1, data.size(), f);
	if (f != stdout) fclose(f);
	return 0;
}
