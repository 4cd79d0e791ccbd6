"""clock64 timeline of one chain-kernel CTA (needs the -DSIMBA_TRAIN_TIMELINE build)."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ['SIMBA_B200_LIB'] = os.path.join(ROOT, 'ethz-safe-learning_b200', 'simba_b200', 'libsimba_b200_tl.so')
sys.argv = [sys.argv[0], '--steps', '20']
exec(open(os.path.join(ROOT, 'tools', 'profile_train.py')).read())
lib = _lib.load()
buf = (C.c_longlong * 512)()
lib.simba_debug_train_timeline.argtypes = [C.c_void_p]
assert lib.simba_debug_train_timeline(buf) == 0
t0 = buf[0]
print("start->loop %d, loop_end(sum) %d" % (buf[1] - t0, buf[2] - t0))
for it in range(9):
    b = [buf[8 + it * 8 + j] - t0 for j in range(5)]
    print("chunk %d: top %6d issue %5d wait %5d fma %5d epi %5d" % (it, b[0], b[1] - b[0], b[2] - b[1], b[3] - b[2], b[4] - b[3]))
