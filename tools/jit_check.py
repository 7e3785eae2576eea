"""Development aid: compile the k_scan_jit source with NVRTC (no GPU needed) and print ptxas statistics."""
import re
import sys
from cuda.bindings import nvrtc

src = open("fastplong_b200/csrc/fpl_scan_jit_src.h").read()
body = re.search(r'R"JITSRC\((.*)\)JITSRC"', src, re.S).group(1)
a0 = sys.argv[1] if len(sys.argv) > 1 else "AATGTACTTCGTTCAGTTACGTATTGCTAA"
a1 = sys.argv[2] if len(sys.argv) > 2 else "TTAGCAATACGTAACTGAACGAAGTACATT"
defs = f'#define FPL_A0 "{a0}"\n#define FPL_A1 "{a1}"\n#define FPL_DO_ADAPTERS true\n#define FPL_DO_COUNTS true\n#define FPL_DO_CPLX false\n#define FPL_QQ 48\n#define FPL_MINBLOCKS 8\n'
err, prog = nvrtc.nvrtcCreateProgram((defs + body).encode(), b"fpl_scan_jit.cu", 0, [], [])
opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-lineinfo", b"--ptxas-options=-v"]
err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
_, n = nvrtc.nvrtcGetProgramLogSize(prog)
log = b" " * n
nvrtc.nvrtcGetProgramLog(prog, log)
print(log.decode(errors="replace")[-3000:])
print("compile status", err)
if int(err) == 0:
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    cubin = b" " * n
    nvrtc.nvrtcGetCUBIN(prog, cubin)
    open("/tmp/scan_jit.cubin", "wb").write(cubin)
    print("cubin bytes", n)
