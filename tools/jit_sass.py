"""Compile a circuit-specialised kernel here (NVRTC cross-compiles for sm_100a without a GPU) and
report registers / spills / SASS size / instruction mix.
    python tools/jit_sass.py c4 R G [ilp] [threads] [minb] [--dump file.sass]"""
import ctypes as C
import glob
import os
import subprocess
import sys
import tempfile
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
name = sys.argv[1]; R = int(sys.argv[2]); G = int(sys.argv[3])
ilp = sys.argv[4] if len(sys.argv) > 4 and not sys.argv[4].startswith('--') else None
thr = sys.argv[5] if len(sys.argv) > 5 and not sys.argv[5].startswith('--') else None
minb = sys.argv[6] if len(sys.argv) > 6 and not sys.argv[6].startswith('--') else None
dump = sys.argv[sys.argv.index('--dump') + 1] if '--dump' in sys.argv else None
cache = tempfile.mkdtemp(prefix='ahk_sass_')
os.environ['AHK_JIT_CACHE'] = cache
if ilp: os.environ['AHK_JIT_ILP'] = ilp
if thr: os.environ['AHK_JIT_THREADS'] = thr
if minb: os.environ['AHK_JIT_MINB'] = minb
from ahkab_b200 import _abi, workloads as W
from ahkab_b200.circuit import Circuit
import oracle
w = W.ALL[name](B=16)
o = oracle.Oracle(w.circuit(Circuit), w.batch(), w.opts)
lib = _abi.load_library()
kind = {'c2': 0, 'c3': 1, 'c4': 1, 'c5': 0}[name]
method, usc = {'c3': ('TRAP', 0), 'c4': ('GEAR2', 1)}.get(name, ('TRAP', 0))
nb = C.c_int64(0)
rc = lib.ahk_jit_compile(C.byref(o.desc.c), C.byref(o.opts), len(o.slots),
                         o.slots.ctypes.data_as(C.POINTER(C.c_int32)), kind, _abi.METHODS[method], usc,
                         1000 + R * 64 + G, C.byref(nb))
if rc != 0:
    print(lib.ahk_last_error().decode()[:4000]); sys.exit(1)
cubin = glob.glob(os.path.join(cache, '*.cubin'))[0]
res = subprocess.run(['cuobjdump', '-res-usage', cubin], capture_output=True, text=True).stdout
print([l for l in res.splitlines() if 'REG' in l][0].strip())
sass = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True).stdout
ops = collections.Counter()
n = 0
for l in sass.splitlines():
    l = l.strip()
    if not l.startswith('/*') or ';' not in l: continue
    body = l.split('*/', 1)[1].strip()
    if body.startswith('@'): body = body.split(' ', 1)[1].strip()
    op = body.split(' ')[0].split('.')[0].rstrip(';')
    ops[op] += 1; n += 1
print('SASS instructions:', n)
print(' '.join('%s:%d' % kv for kv in ops.most_common(28)))
if dump:
    open(dump, 'w').write(sass)
