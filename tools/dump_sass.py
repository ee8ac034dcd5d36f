"""Compile one model/method/output variant with NVRTC (no GPU needed) and write the cubin, so that
`cuobjdump -sass` / `nvdisasm` can be run on it.  usage:
  python tools/dump_sass.py strains discrete_time_stochastic stats 32 256 4 /tmp/k.cubin [inject] [--flags N]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples import configs, models
from popdynamics_b200 import capi, compiler

MODELS = {'strains': models.StrainsModel, 'sir': models.SirModel, 'seir': lambda: configs.seir_sweep(0, 8),
          'tb': models.SimpleTbModel, 'tbsu': models.ScaleupTbModel, 'rmit': lambda: configs.rmit_sweep(0, 8)}


def main():
    name, method, out, bits, block, minb, path = sys.argv[1:8]
    rng = capi.INJECT if len(sys.argv) > 8 and sys.argv[8] == 'inject' else capi.PHILOX
    model = MODELS[name]()
    model.make_times(0, 10, 1)
    model.check_flow_transfers()
    cm = compiler.compile_model(model, method)
    if '--header' in sys.argv:
        print(cm.cuda_source())
    k = capi.Model(cm.cuda_source(), capi.METHOD_CODES[method], capi.OUTPUT_CODES[out], rng, 0,
                   cm.n_comp, cm.n_param, len(cm.member_param_slots()), cm.n_su, cm.n_flow,
                   count_bits=int(bits), block=int(block), min_blocks=int(minb),
                   flags=int(sys.argv[sys.argv.index('--flags') + 1]) if '--flags' in sys.argv else 0)
    n = C.c_size_t()
    capi.check(capi.lib().pd_model_cubin_size(k.handle, C.byref(n)))
    buf = C.create_string_buffer(n.value)
    capi.check(capi.lib().pd_model_cubin_copy(k.handle, buf, n))
    open(path, 'wb').write(buf.raw)
    print('wrote', path, n.value, 'bytes')


if __name__ == '__main__':
    main()
