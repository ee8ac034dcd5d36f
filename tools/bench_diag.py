"""Throughput of the batched diagnostics pass (pd_diag_kernel): elementwise, HBM-bound.
usage: python tools/bench_diag.py [members] [n_time]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from examples import configs
from popdynamics_b200 import ensemble


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 101
    model = configs.strains()
    model.make_times(0, T - 1, 1)
    C = len(model.labels)
    res = ensemble.EnsembleResult()
    res.method, res.labels, res.n_members, res.device = 'discrete_time_stochastic', \
        list(model.labels), M, 0
    res.times = torch.arange(T, dtype=torch.float64).numpy()
    res.soln = torch.randint(1, 2000, (T, C, M), device='cuda').to(torch.float64)
    model.check_flow_transfers()
    for want_flows in (False, True):
        best, d = 1e30, None
        for _ in range(3):
            d = ensemble.diagnostics(model, res, want_flows=want_flows)
            best = min(best, d.info['kernel_ms'])
        V = len(d.var_labels)
        bytes_moved = 8.0 * M * T * (C + V + (C if want_flows else 0))
        print(json.dumps({'kernel': 'pd_diag_kernel', 'members': M, 'n_time': T, 'vars': V,
                          'want_flows': want_flows, 'kernel_ms': best,
                          'algorithmic_GBps': bytes_moved / (best * 1e-3) / 1e9,
                          'member_times_per_s': M * T / (best * 1e-3)}))
        del d
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
