#!/bin/bash
# Host-parse A/B on the GPU box's cores (this container's timings wander by 20 %):
#   tools/parse_ab.sh c1|c3|c4 binary [binary ...]      binaries = parsebench builds under tools/_bin
# prints each binary's best ms per picture over five alternating rounds.
cd "$(dirname "$0")/.."
case $1 in
  c1) codec=h264;  tools/_bin/streamgen h264 mbw=120 mbh=68 frames=16 seed=100 crop_bottom=8 gop=1000 refs=1 qp=28 intra_pct=10 skip_pct=15 coded_pct=25 deblock=1 > /tmp/ab.es;;
  c3) codec=h264;  tools/_bin/streamgen h264 mbw=120 mbh=68 frames=16 seed=100 crop_bottom=8 refs=2 bframes=2 cabac=1 profile=100 transform8x8=1 weighted_pred=1 weighted_bipred=1 gop=1000 qp=28 intra_pct=10 skip_pct=15 coded_pct=25 > /tmp/ab.es;;
  c4) codec=mpeg2; tools/_bin/streamgen mpeg2 mbw=120 mbh=68 frames=12 seed=100 bframes=2 gop=12 mv_range=64 > /tmp/ab.es;;
  *) echo "usage: parse_ab.sh c1|c3|c4 binary..."; exit 2;;
esac
shift
for round in 1 2 3 4 5; do
  for b in "$@"; do
    echo "$b $(B2_PARSE_AS_DEVICE=1 tools/_bin/$b $codec /tmp/ab.es 15 | sed 's/.*pictures, \([0-9.]*\) ms.*/\1/')"
  done
done | awk '{ if (!($1 in m) || $2 < m[$1]) m[$1] = $2 } END { for (k in m) printf "%s %.3f ms per picture\n", k, m[k] }' | sort
