"""dev experiment: do two half-size pipelines on two streams overlap into less time than one full-size pipeline?
(the kernels are bound by different things: pre-tokeniser = ALU pipe, probe / miss = L2 latency)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B

K = 6
half = 512 << 20
full = B.Bench("config2", 2 * half, 0, 1, 0)
a = B.Bench("config2", half, 0, 1, 0, seed_offset=1)
b = B.Bench("config2", half, 0, 1, 0, seed_offset=2)
c = B.Bench("config2", half // 4, 0, 1, 0, seed_offset=3)          # phase shifter
for x in (full, a, b, c):
    for _ in range(2):
        x.step_sync()

def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / K

def run_full():
    for _ in range(K): full.enqueue()
    full.core.device_wait()
def run_seq():
    for _ in range(K): a.enqueue(); a.core.device_wait(); b.enqueue(); b.core.device_wait()
def run_par(shift):
    if shift: c.enqueue()
    for _ in range(K): a.enqueue(); b.enqueue()
    a.core.device_wait(); b.core.device_wait()
    if shift: c.core.device_wait()
b_stream = b.stream
print("full 1 GiB, one stream      : %.2f ms/step" % timed(run_full))
print("two halves, serial          : %.2f ms/step" % timed(run_seq))
print("two halves, two streams     : %.2f ms/step" % timed(lambda: run_par(False)))
c.stream = b.stream                                                  # the shifter runs ahead of b on b's stream
print("two halves, two streams, b shifted by a quarter-size job: %.2f ms/step (incl. the shifter once)" % timed(lambda: run_par(True)))
