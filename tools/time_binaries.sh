#!/bin/bash
# Development aid (run under gpurun): wall time of the reference binary and of the GPU drop-in on BASELINE config 1.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import sys; sys.path.insert(0, '.')
from fastplong_b200 import synth
b = synth.ont_like(10000, 8000, 1)
synth.to_fastq(b, '/dev/shm/c1.fq')
print('bases', b.n_bases)
PY
S=AATGTACTTCGTTCAGTTACGTATTGCTAA
TIMEFORMAT="%R s"
for w in 16; do echo -n "ref -w $w: "; { time oracle/_ref/fastplong_ref -i /dev/shm/c1.fq -o /dev/shm/ref.fq -s $S -w $w -j /dev/shm/r.json -h /dev/shm/r.html >/dev/null 2>&1; } 2>&1; done
for w in 1 4 8; do echo -n "gpu -w $w: "; { time build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w $w -j /dev/shm/g.json -h /dev/shm/g.html -V 2>/dev/shm/g.err >/dev/null; } 2>&1; grep -E "start to|Loading completed|writer finished" /dev/shm/g.err | tr '\n' ' '; echo; done
cmp /dev/shm/ref.fq /dev/shm/gpu.fq && echo "outputs identical"
rm -f /dev/shm/c1.fq /dev/shm/ref.fq /dev/shm/gpu.fq
