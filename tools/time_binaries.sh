#!/bin/bash
# Development aid (run under gpurun): wall time of the reference binary and of the GPU drop-in on a config-1 shaped
# FASTQ file (N reads generated, then the file is concatenated REP times).  usage: tools/time_binaries.sh [N] [REP]
cd "$(dirname "$0")/.."
N=${1:-50000}; REP=${2:-1}
python - <<PY
import sys; sys.path.insert(0, '.')
from fastplong_b200 import synth
b = synth.ont_like($N, 8000, 1)
synth.to_fastq(b, '/dev/shm/c1_part.fq')
print('reads', b.n_reads * $REP, 'bases', b.n_bases * $REP)
PY
rm -f /dev/shm/c1.fq; for i in $(seq $REP); do cat /dev/shm/c1_part.fq >> /dev/shm/c1.fq; done; rm /dev/shm/c1_part.fq
ls -la /dev/shm/c1.fq
S=AATGTACTTCGTTCAGTTACGTATTGCTAA
TIMEFORMAT="%R s"
build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 2 -j /dev/shm/g.json -h /dev/shm/g.html --reads_to_process 1000 >/dev/null 2>&1   # warm the image
for rep in 1 2; do
echo -n "ref -w 16: "; { time oracle/_ref/fastplong_ref -i /dev/shm/c1.fq -o /dev/shm/ref.fq -s $S -w 16 -j /dev/shm/r.json -h /dev/shm/r.html >/dev/null 2>&1; } 2>&1
echo -n "gpu -w 4 device parse: "; { time build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 4 -j /dev/shm/g.json -h /dev/shm/g.html >/dev/null 2>&1; } 2>&1
done
cmp /dev/shm/ref.fq /dev/shm/gpu.fq && echo "outputs identical"
FPL_TIMING=1 build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 4 -j /dev/shm/g.json -h /dev/shm/g.html 2>&1 | grep -E "fastplong_gpu\]|libfplgpu\]"
echo -n "gpu -w 4 reference reader: "; { time FPL_HOST_PARSE=1 build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 4 -j /dev/shm/g.json -h /dev/shm/g.html >/dev/null 2>&1; } 2>&1
rm -f /dev/shm/c1.fq /dev/shm/ref.fq /dev/shm/gpu.fq
