#!/bin/bash
# Development aid (run under gpurun): where the drop-in CLI's wall time goes on a 6.4 GB FASTQ, with different writer
# backlogs.  usage: tools/cli_probe.sh
cd "$(dirname "$0")/.."
python - <<PY
import sys; sys.path.insert(0, '.')
from fastplong_b200 import synth
b = synth.ont_like(50000, 8000, 1)
synth.to_fastq(b, '/dev/shm/c1_part.fq')
PY
rm -f /dev/shm/c1.fq; for i in $(seq 8); do cat /dev/shm/c1_part.fq >> /dev/shm/c1.fq; done; rm /dev/shm/c1_part.fq
S=AATGTACTTCGTTCAGTTACGTATTGCTAA
build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 2 -j /dev/shm/g.json -h /dev/shm/g.html --reads_to_process 1000 >/dev/null 2>&1
for bl in 2 8 1000; do
  echo "== FPL_WRITER_BACKLOG=$bl"
  FPL_WRITER_BACKLOG=$bl FPL_TIMING=1 build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 4 -j /dev/shm/g.json -h /dev/shm/g.html 2>&1 | grep -E "fastplong_gpu\]"
done
echo "== FPL_JIT_V1=1"
FPL_JIT_V1=1 FPL_TIMING=1 build/fastplong_gpu -i /dev/shm/c1.fq -o /dev/shm/gpu.fq -s $S -w 4 -j /dev/shm/g.json -h /dev/shm/g.html 2>&1 | grep -E "fastplong_gpu\]|specialisation"
rm -f /dev/shm/c1.fq /dev/shm/gpu.fq
