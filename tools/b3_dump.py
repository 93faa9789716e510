"""Dump the generated source of the compiled C5 elimination schedule (biglin3_gen.hpp) and, with
--ptxas, compile it with nvcc for sm_100a to see registers / spills: python tools/b3_dump.py [nodes] [--ptxas]"""
import ctypes as C
import subprocess
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/hostemu')
import emu
from ahkab_b200 import workloads as W
from ahkab_b200.circuit import Circuit

nodes = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
w = W.c5_r2r(6, nodes=nodes)
e = emu.Emu(w.circuit(Circuit), w.batch(), w.opts)
lib = emu.lib()
c6 = e._common()
nf, npos = C.c_int32(0), C.c_int32(0)
args = (c6[0], c6[1], c6[3], c6[4])
n = lib.emu_big3_gen_source(*args, None, C.c_int64(0), C.byref(nf), C.byref(npos))
buf = C.create_string_buffer(n + 1)
lib.emu_big3_gen_source(*args, buf, C.c_int64(n + 1), C.byref(nf), C.byref(npos))
open('/tmp/b3.cu', 'wb').write(buf.value)
print('source bytes', n, 'nf', nf.value, 'npos', npos.value)
if '--ptxas' in sys.argv:
    subprocess.call('/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -lineinfo '
                    '-Xptxas -v -cubin -o /tmp/b3.cubin /tmp/b3.cu', shell=True)
