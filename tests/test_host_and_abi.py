"""CPU tests (no GPU needed): host-side parameter packing, the C-ABI library's exported symbols,
struct layout agreement and the loud failure without a CUDA device."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from ast_sac_b200 import _lib as L
from ast_sac_b200 import env as E
from ast_sac_b200 import scenarios as S
from oracle import oracle as O

from helpers import golden, golden_names
from product_helpers import assets_from_meta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _struct_bytes(s):
    return np.frombuffer(C.string_at(C.byref(s), C.sizeof(s)), dtype=np.uint8)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "shipenv.h")).read()
    declared = sorted(set(re.findall(r"\b(shipenv_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), f"libshipenv.so does not export {name}"
    assert sorted(L.EXPORTS) == declared


def test_struct_layout_matches_header():
    lib = L.load()
    assert lib.shipenv_abi_version() == L.ABI_VERSION
    assert lib.shipenv_sizeof_params() == C.sizeof(L.Params)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    args = S.get_env_args()
    assets, m = S.build_rl_assets(args)
    P = E.pack_params(assets, m, args, L.ENV_RL, post_reset=True)
    h = C.c_void_p()
    rc = L.load().shipenv_create(C.byref(P), 4, 0, C.byref(h))
    assert rc == 2 and not h.value
    assert b"no CPU fallback" in L.load().shipenv_last_error()
    with pytest.raises(RuntimeError, match="CUDA"):
        S.prepare_multiship_rl_env(args)


def test_create_validates_arguments():
    args = S.get_env_args()
    assets, m = S.build_rl_assets(args)
    P = E.pack_params(assets, m, args, L.ENV_RL, post_reset=True)
    h = C.c_void_p()
    lib = L.load()
    assert lib.shipenv_create(C.byref(P), 0, 0, C.byref(h)) == 1          # num_envs
    P.abi_version = 99
    assert lib.shipenv_create(C.byref(P), 4, 0, C.byref(h)) == 1
    assert b"abi_version" in lib.shipenv_last_error()
    P.abi_version = L.ABI_VERSION
    P.ship[1].model_kind = L.MODEL_SIMPLE
    assert lib.shipenv_create(C.byref(P), 4, 0, C.byref(h)) == 1
    assert lib.shipenv_create(None, 4, 0, C.byref(h)) == 1


@pytest.mark.parametrize("name", golden_names("rl_") + golden_names("colav_"))
def test_host_objects_reproduce_reference_config(name):
    """The oracle config extracted from this package's host objects is byte-identical to the one
    extracted from the reference's own objects when the golden fixture was made."""
    g = golden(name)
    meta = json.loads(str(g["meta"]))
    assets, m, args = assets_from_meta(meta)
    kind = {"rl": O.ENV_RL, "colav": O.ENV_COLAV_IW, "noniw": O.ENV_COLAV_NONIW, "noniw_step": O.ENV_COLAV_NONIW}[meta["kind"]]
    cfg = O.env_config_from_assets(assets, m, args, kind)
    assert np.array_equal(_struct_bytes(cfg), g["cfg"])


def test_pack_params_derived_constants():
    args = S.get_env_args(time_step=4)
    assets, m = S.build_rl_assets(args)
    P = E.pack_params(assets, m, args, L.ENV_RL, post_reset=False)
    t = P.ship[0]
    # BaseShipModel.__init__ (ship_model.py:70-78) with the run/env_setup.py numbers
    mass = (3850000 / 0.7 - 3850000) + 0.9 * (3850000 - 200000) + 200000 + 200000
    assert t.mass == mass and t.i_z == mass * (80 ** 2 + 16 ** 2) / 12
    assert t.x_du == mass * 0.4 and t.inv_m_u == 1.0 / (mass + mass * 0.4)
    assert t.dt == 4 and t.dt_shaft == 4 and t.ctrl_dt == 4
    # PTI: ME 0 W, electrical 2 x 510 kW - 200 kW hotel load (ship_engine.py:33-36)
    assert t.p_me == 0 and t.p_el == 2 * 510e3 - 200000
    assert t.thrust_coeff == 3.1 ** 4 * 1.7 and t.init_omega == 420 * np.pi / 30
    assert t.nav_fail_tol == 3000 and P.ship[1].nav_fail_tol == 500
    assert P.n_poly == 6 and P.poly_start[6] == 55
    assert (P.map_min_n, P.map_max_n, P.map_min_e, P.map_max_e) == (0, 10000, 0, 20000)
    # init_get_intermediate_waypoints (env.py:143-162) for the (10000,15000)->(0,5000) route, 9 samplings
    assert abs(P.ab_segment_length - np.hypot(10000, 10000) / 10) < 1e-9
    assert P.n_base0 == 9000 and P.e_base0 == 14000
    # reset() quirk: shaft integrator step becomes 0.01 (ship_engine.py:331-333)
    P2 = E.pack_params(assets, m, args, L.ENV_RL, post_reset=True)
    assert P2.ship[0].dt_shaft == 0.01 and P2.ship[1].dt_shaft == 0.01


def test_pack_params_collav_modes_and_bad_routes():
    for mode, code in (('none', L.COLLAV_NONE), ('simple', L.COLLAV_SIMPLE), ('sbmpc', L.COLLAV_SBMPC)):
        args = S.get_env_args(collav_mode=mode)
        assets, m = S.build_rl_assets(args)
        P = E.pack_params(assets, m, args, L.ENV_RL, post_reset=True)
        assert P.collav == code
        assert P.ship[1].w_ship == assets[1].ship_model.w_ship and P.ship[1].l_ship == assets[1].ship_model.l_ship
    args = S.get_env_args(collav_mode='mpc')
    with pytest.raises(ValueError):
        E.pack_params(assets, m, args, L.ENV_RL, post_reset=True)
    with pytest.raises((OSError, FileNotFoundError)):      # same exception type as the reference's np.loadtxt
        from ast_sac_b200.sim.LOS_guidance import NavigationSystem
        NavigationSystem("/nonexistent/route.txt")


def test_event_strings_match_reference_order():
    assert E.events_to_string(1) == 'Ships collision!'
    assert E.events_to_string((1 << 5) | (1 << 7)) == ('|Ship under test reaches its final destination!|'
                                                      '|Obstacle ship reaches its final destination!|')
    assert E.EVENT_STRINGS == O.EVENT_STRINGS


def test_layout_constants_match_header():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "shipenv.h")).read()
    assert int(re.search(r"#define SHIPENV_ABI_VERSION (\d+)", hdr).group(1)) == L.ABI_VERSION
    sf = re.search(r"/\* ship_f64 rows \*/\s*enum \{(.*?)SHIPENV_SF_COUNT", hdr, re.S).group(1)
    assert len(re.findall(r"SHIPENV_SF_\w+", sf)) == L.SF_COUNT == len(L.SF)
    ef = re.search(r"/\* env_f64 rows \*/\s*enum \{(.*?)SHIPENV_EF_COUNT", hdr, re.S).group(1)
    assert len(re.findall(r"SHIPENV_EF_\w+", ef)) == L.EF_COUNT == len(L.EF)
    lg = re.search(r"SHIPENV_LOG_TIME = 0,(.*?)SHIPENV_LOG_COLS", hdr, re.S).group(1)
    assert 1 + len(re.findall(r"SHIPENV_LOG_\w+", lg)) == len(L.LOG_COLS)


def test_batched_info_decodes_lazily_and_like_the_eager_dict():
    """BatchedInfo (env_info of a batched call): the reference's keys, decoded from the packed info words on first
    access and cached; a read-only Mapping."""
    import torch
    words = torch.tensor([0, 0x3 | L.INFO_TERMINAL | L.INFO_DONE, 0x40 | L.INFO_TEST_STOP,
                          0x100 | L.INFO_OBS_STOP | L.INFO_TEST_STOP | L.INFO_DONE], dtype=torch.int32)
    nsub = torch.tensor([5, 7, 0, 9], dtype=torch.int32)
    info = E.BatchedInfo(words, nsub)
    assert list(info) == ['events', 'terminal', 'test_ship_stop', 'obs_ship_stop', 'substeps'] and len(info) == 5
    assert info._cache == {}                                          # nothing decoded yet
    assert info['events'].tolist() == [0, 0x3, 0x40, 0x100]
    assert info['terminal'].tolist() == [False, True, False, False]
    assert info['test_ship_stop'].tolist() == [False, False, True, True]
    assert info['obs_ship_stop'].tolist() == [False, False, False, True]
    assert info['substeps'] is nsub
    assert info['terminal'] is info['terminal']                        # cached
    assert dict(info).keys() == {'events', 'terminal', 'test_ship_stop', 'obs_ship_stop', 'substeps'}
    assert info.get('missing') is None
    with pytest.raises(KeyError):
        info['missing']
    with pytest.raises(TypeError):
        info['terminal'] = 1                                           # Mapping, not MutableMapping
