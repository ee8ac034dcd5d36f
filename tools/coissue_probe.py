"""Co-issue probes (pd_microbench kinds 6-10, csrc/popdyn_capi.cu): does an instruction of another
pipe hide behind a DADD (two cycles of the FP64 pipe) or an IMAD.WIDE (four of the FMA-heavy pipe),
or do issue slots add up?  Prints G elements/s and scheduler cycles per warp-element.
usage: python tools/coissue_probe.py   (needs cuda:0)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from popdynamics_b200 import capi

WHAT = {0: 'DMUL + DADD per element (2 fp64 instructions)', 1: 'DFMA', 2: 'IMAD.WIDE + LOP3',
        3: 'IMAD.HI + LOP3', 4: 'IMAD (low half) + LOP3', 5: 'SHF + SHF + LOP3',
        6: 'DADD alone', 7: 'DADD + 1 LOP3 on its own chain', 8: 'DADD + 2 LOP3',
        9: 'IMAD.WIDE + LOP3 + 1 LOP3 on its own chain', 10: 'IMAD.WIDE + LOP3 + 2 LOP3'}


def main():
    props = torch.cuda.get_device_properties(0)
    schedulers = props.multi_processor_count * 4
    mhz = float(os.environ.get('PD_SM_MHZ', 1965.0))
    for kind in range(11):
        gops = capi.microbench(kind, 0)
        per_elem = {0: 2.0, 1: 1.0, 2: 3.0}.get(kind, 1.0)
        elems = gops / per_elem                       # G elements/s
        cycles = mhz * 1e6 / (elems * 1e9 / 32 / schedulers)
        print(json.dumps({'kind': kind, 'what': WHAT[kind], 'g_elements_per_s': round(elems, 1),
                          'cycles_per_warp_element': round(cycles, 3),
                          'assumed_sm_mhz': mhz}))


if __name__ == '__main__':
    main()
