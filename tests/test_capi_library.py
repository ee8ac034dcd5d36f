"""The C-ABI library loads without a GPU and exports every symbol include/popdyn_b200.h declares;
host-side entry points work; run entry points fail loudly (no CPU fallback) without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle
from popdynamics_b200 import capi

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'popdyn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_all_exported(built_library):
    lib = ctypes.CDLL(built_library)
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), 'missing export %s' % name
    assert sorted(capi.EXPORTED_SYMBOLS) == names


def test_header_constants_match_the_binding():
    text = open(os.path.join(ROOT, 'include', 'popdyn_b200.h')).read()
    defines = dict(re.findall(r'#define\s+(PD_[A-Z_0-9]+)\s+(-?\d+)\b', text))
    assert int(defines['PD_FLAG_MEMBER_MASK']) == capi.FLAG_MEMBER_MASK
    assert int(defines['PD_FLAG_HIST_MOMENTS']) == capi.FLAG_HIST_MOMENTS
    enums = dict(re.findall(r'\b(PD_[A-Z_]+)\s*=\s*(-?\d+)', text))
    assert int(enums['PD_DIAGNOSTICS']) == capi.DIAGNOSTICS
    assert int(enums['PD_ADAPTIVE']) == capi.ADAPTIVE == capi.METHOD_CODES['scipy']
    assert (int(enums['PD_EXPLICIT']), int(enums['PD_TAU_LEAP']), int(enums['PD_GILLESPIE'])) == \
        (capi.EXPLICIT, capi.TAU_LEAP, capi.GILLESPIE)
    for name, value in capi.METHOD_CODES.items():
        assert value in (0, 1, 2, 4)
    assert int(enums['PD_E_INVALID']) < 0 and int(enums['PD_E_NO_DEVICE']) == capi.PD_E_NO_DEVICE


def test_version_and_device_count(built_library):
    assert capi.lib().pd_version() == 100
    assert capi.device_count() >= 0


def test_native_schedule_matches_oracle(built_library):
    m = helpers.make_model('tb')
    m.make_times(1900, 1910, .05)
    for tt in (list(range(0, 201)), m.target_times, [0., 0.3, 0.31, 5.]):
        t_eval, dt, n_sub = capi.explicit_schedule(tt)
        o_eval, o_dt, o_interval = oracle.explicit_schedule(tt)
        assert np.array_equal(t_eval, o_eval) and np.array_equal(dt, o_dt)
        assert np.array_equal(np.repeat(np.arange(1, len(tt)), n_sub), o_interval)


def test_structs_match_the_compiled_layout(built_library):
    size = capi.lib().pd_struct_size
    assert ctypes.sizeof(capi.ModelDesc) == size(0)
    assert ctypes.sizeof(capi.RunResult) == size(1)
    assert ctypes.sizeof(capi.StochRun) == size(2)
    assert ctypes.sizeof(capi.ExplicitRun) == size(3)
    assert ctypes.sizeof(capi.DiagRun) == size(4)
    assert ctypes.sizeof(capi.AdaptiveRun) == size(5)
    assert size(6) == -1


@pytest.mark.skipif(capi.lib().pd_device_count() > 0, reason='a GPU is present')
def test_no_cpu_fallback_without_a_gpu(built_library):
    m = helpers.make_model('sir')
    m.make_times(0, 5, 1)
    with pytest.raises(capi.PopdynError) as exc:
        m.integrate()
    assert exc.value.status in (capi.PD_E_NO_DEVICE, -3)
    with pytest.raises(capi.PopdynError):
        m.integrate_ensemble('discrete_time_stochastic', n_members=4)


def test_missing_extension_is_loud(monkeypatch):
    monkeypatch.setattr(capi, '_lib', None)
    monkeypatch.setattr(capi, 'LIB_PATH', '/nonexistent/libpopdyn_b200.so')
    with pytest.raises(capi.ExtensionMissing):
        capi.lib()
