#!/bin/bash
# Development aid (run under gpurun): where the drop-in CLI's wall time goes on a 6.4 GB FASTQ (400 k reads of the
# config-1 generator, files on tmpfs), host-side and device-side output assembly, against the reference binary.
# usage: tools/cli_probe.sh [reference-too]
cd "$(dirname "$0")/.."
python - <<PY
import sys; sys.path.insert(0, '.')
from fastplong_b200 import synth
b = synth.ont_like(50000, 8000, 1)
synth.to_fastq(b, '/dev/shm/c1_part.fq')
PY
rm -f /dev/shm/c1.fq; for i in $(seq 8); do cat /dev/shm/c1_part.fq >> /dev/shm/c1.fq; done; rm /dev/shm/c1_part.fq
S=AATGTACTTCGTTCAGTTACGTATTGCTAA
ARGS="-i /dev/shm/c1.fq -s $S -j /dev/shm/g.json -h /dev/shm/g.html"
build/fastplong_gpu $ARGS -o /dev/shm/gpu.fq -w 2 --reads_to_process 1000 >/dev/null 2>&1      # page the binary and the driver in
t() { local a=$(date +%s%N); "$@" >/dev/null 2>&1; local b=$(date +%s%N); echo "$(( (b - a) / 1000000 )) m"; }
for rep in 1 2 3 4; do
  rm -f /dev/shm/gpu.fq; echo "gpu -w 4, host assembly:   $(t build/fastplong_gpu $ARGS -o /dev/shm/gpu.fq -w 4) s"
  rm -f /dev/shm/gpu2.fq; echo "gpu -w 4, device assembly: $(FPL_DEVICE_EMIT=1 t build/fastplong_gpu $ARGS -o /dev/shm/gpu2.fq -w 4) s"
done
cmp /dev/shm/gpu.fq /dev/shm/gpu2.fq && echo "host- and device-assembled outputs identical"
echo "== stages, host assembly"
FPL_TIMING=1 build/fastplong_gpu $ARGS -o /dev/shm/gpu.fq -w 4 2>&1 | grep -E "fastplong_gpu\]"
echo "== stages, device assembly"
FPL_DEVICE_EMIT=1 FPL_TIMING=1 build/fastplong_gpu $ARGS -o /dev/shm/gpu2.fq -w 4 2>&1 | grep -E "fastplong_gpu\]"
if [ -n "$1" ]; then
  rm -f /dev/shm/ref.fq; echo "ref -w 16: $(t oracle/_ref/fastplong_ref $ARGS -o /dev/shm/ref.fq -w 16) s"
  cmp /dev/shm/ref.fq /dev/shm/gpu.fq && echo "reference and gpu outputs identical"
fi
rm -f /dev/shm/c1.fq /dev/shm/gpu.fq /dev/shm/gpu2.fq /dev/shm/ref.fq
