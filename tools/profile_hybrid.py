"""Host-side profile of ensemble.integrate_hybrid on the mix_model workload (bench config 5b)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from examples import configs
from popdynamics_b200 import ensemble


def run(M):
    model = configs.mix_model()
    t0 = time.perf_counter()
    res = ensemble.integrate_hybrid(model, configs.MIX_SEGMENTS, M, configs.MIX_CUTOFF, seed=5)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, res


def main():
    big = torch.empty(5 << 30, dtype=torch.uint8, device='cuda')
    del big
    torch.cuda.empty_cache()
    print('first (compiles)', run(4096)[0])
    for _ in range(3):
        print('2^20 members', run(1 << 20)[0])
    prof = cProfile.Profile()
    prof.enable()
    run(1 << 20)
    prof.disable()
    pstats.Stats(prof).sort_stats('cumulative').print_stats(25)


if __name__ == '__main__':
    main()
