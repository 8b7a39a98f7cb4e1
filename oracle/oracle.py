"""ctypes wrapper around ``oracle/mate_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

See the header of ``mate_oracle.c``.  The POD structs are the public ABI structs of
``include/mate_b200.h`` (mirrored in ``mate_b200/_abi.py``) so that the same state /
aux / config objects can be handed to the oracle and to the CUDA library.
"""

import ctypes
import os
import subprocess

import numpy as np

from mate_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libmate_oracle.so')


def build(force=False):
    src = os.path.join(HERE, 'mate_oracle.c')
    hdr = os.path.join(HERE, '..', 'include', 'mate_b200.h')
    stale = (
        force
        or not os.path.exists(LIB)
        or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.run(['make', '-C', HERE, '-B'], check=True, capture_output=True)
    return LIB


_lib = None


def _load():
    global _lib  # pylint: disable=global-statement
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        vp = ctypes.c_void_p
        lib.oracle_create.restype = vp
        lib.oracle_create.argtypes = [ctypes.POINTER(_abi.MateConfig), ctypes.c_int32, ctypes.c_int64]
        lib.oracle_destroy.argtypes = [vp]
        lib.oracle_set_state.argtypes = [vp, ctypes.POINTER(_abi.MateStateView)]
        lib.oracle_get_state.argtypes = [vp, ctypes.POINTER(_abi.MateStateView)]
        lib.oracle_get_fov.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int]
        lib.oracle_observe.argtypes = [vp, vp, ctypes.c_uint64, vp, vp, ctypes.POINTER(_abi.MateStepAux)]
        lib.oracle_reset.argtypes = [vp, vp, ctypes.c_uint64, vp, vp]
        lib.oracle_step.argtypes = [vp, vp, vp, vp, vp, ctypes.c_uint64, ctypes.c_uint32, vp, vp, vp, vp,
                                    ctypes.POINTER(_abi.MateStepAux)]
        lib.oracle_episode_stats.argtypes = [vp, vp, ctypes.c_int]
        lib.oracle_set_threads.argtypes = [vp, ctypes.c_int]
        lib.oracle_obs_dims.argtypes = [vp, _abi.c_int32_p, _abi.c_int32_p]
        _lib = lib
    return _lib


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """Batched float64 CPU oracle with the same call surface as the CUDA simulator."""

    def __init__(self, flat_config, num_envs, env_index_base=0, num_threads=1):
        self.lib = _load()
        self.cfg = dict(flat_config)
        self._cfg_struct = _abi.make_config_struct(self.cfg)
        self.B = int(num_envs)
        self.nc = int(self.cfg['num_cameras'])
        self.nt = int(self.cfg['num_targets'])
        self.no = int(self.cfg['num_obstacles'])
        self.handle = self.lib.oracle_create(ctypes.byref(self._cfg_struct), self.B, int(env_index_base))
        if not self.handle:
            raise ValueError('oracle_create failed (unsupported configuration)')
        dc, dt = ctypes.c_int32(), ctypes.c_int32()
        self.lib.oracle_obs_dims(self.handle, ctypes.byref(dc), ctypes.byref(dt))
        self.dc, self.dt = dc.value, dt.value
        self.set_threads(num_threads)

    def __del__(self):
        if getattr(self, 'handle', None):
            self.lib.oracle_destroy(self.handle)
            self.handle = None

    def set_threads(self, n):
        self.lib.oracle_set_threads(self.handle, int(n))

    def alloc_aux(self):
        return {n: np.zeros(shape(self.B, self.nc, self.nt, self.no), dtype=dt) for n, _, dt, shape in _abi.AUX_FIELDS}

    @staticmethod
    def _aux_struct(aux):
        if aux is None:
            return None
        s = _abi.MateStepAux()
        for name, ctype, _, _ in _abi.AUX_FIELDS:
            arr = aux.get(name)
            setattr(s, name, ctype() if arr is None else arr.ctypes.data_as(ctype))
        return s

    def set_state(self, arrays):
        arrays = dict(arrays)
        view = _abi.state_view_from_arrays(arrays)
        self.lib.oracle_set_state(self.handle, ctypes.byref(view))

    def get_state(self):
        arrays = _abi.alloc_state_arrays(self.B, self.nc, self.nt, self.no)
        view = _abi.state_view_from_arrays(arrays)
        self.lib.oracle_get_state(self.handle, ctypes.byref(view))
        return arrays

    def get_fov(self, b, c):
        cap = 360 + self.no * 186 + 2
        phi, rho = np.zeros(cap), np.zeros(cap)
        n = self.lib.oracle_get_fov(self.handle, b, c, _ptr(phi), _ptr(rho), cap)
        return phi[:n].copy(), rho[:n].copy()

    def observe(self, transmit=None, seed=0, aux=None):
        cam = np.zeros((self.B, self.nc, self.dc))
        tgt = np.zeros((self.B, self.nt, self.dt))
        transmit = None if transmit is None else np.ascontiguousarray(transmit, dtype=np.uint8)
        s = self._aux_struct(aux)
        self.lib.oracle_observe(self.handle, _ptr(transmit), int(seed), _ptr(cam), _ptr(tgt),
                                None if s is None else ctypes.byref(s))
        return cam, tgt

    def reset(self, seed=0, env_mask=None):
        cam = np.zeros((self.B, self.nc, self.dc))
        tgt = np.zeros((self.B, self.nt, self.dt))
        env_mask = None if env_mask is None else np.ascontiguousarray(env_mask, dtype=np.uint8)
        self.lib.oracle_reset(self.handle, _ptr(env_mask), int(seed), _ptr(cam), _ptr(tgt))
        return cam, tgt

    def step(self, cam_act, tgt_act, transmit=None, goal_choice=None, seed=0, auto_reset=False, aux=None,
             want_obs=True):
        cam_act = np.ascontiguousarray(cam_act, dtype=np.float64).reshape(self.B, self.nc, 2)
        tgt_act = np.ascontiguousarray(tgt_act, dtype=np.float64).reshape(self.B, self.nt, 2)
        transmit = None if transmit is None else np.ascontiguousarray(transmit, dtype=np.uint8)
        goal_choice = None if goal_choice is None else np.ascontiguousarray(goal_choice, dtype=np.int8)
        cam = np.zeros((self.B, self.nc, self.dc)) if want_obs else None
        tgt = np.zeros((self.B, self.nt, self.dt)) if want_obs else None
        rewards = np.zeros((self.B, 2))
        done = np.zeros(self.B, dtype=np.uint8)
        s = self._aux_struct(aux)
        self.lib.oracle_step(self.handle, _ptr(cam_act), _ptr(tgt_act), _ptr(transmit), _ptr(goal_choice),
                             int(seed), _abi.MATE_STEP_AUTO_RESET if auto_reset else 0, _ptr(cam), _ptr(tgt),
                             _ptr(rewards), _ptr(done), None if s is None else ctypes.byref(s))
        return (cam, tgt), rewards, done

    def episode_stats(self, reset_after=False):
        out = np.zeros(16)
        self.lib.oracle_episode_stats(self.handle, _ptr(out), int(reset_after))
        return out
