# host parse alone: one process, then one per core at the same time (memory / cache contention without any GPU work)
cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'.'); import bench; open('/tmp/s.264','wb').write(bench.make_stream(100,16))"
echo "1 process:"; B2_PARSE_AS_DEVICE=1 tools/_bin/parsebench h264 /tmp/s.264 40
N=$(nproc); echo "$N processes at once:"
for i in $(seq $N); do B2_PARSE_AS_DEVICE=1 tools/_bin/parsebench h264 /tmp/s.264 40 > /tmp/pb_$i.log & done; wait
cat /tmp/pb_*.log | awk '{s+=$3; n++} END {print "mean", s/n, "ms per picture over", n, "processes"}'
