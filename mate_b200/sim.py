"""Low-level Python binding of the C ABI (``include/mate_b200.h``).

PyTorch is used only for device memory and streams: tensors are allocated here and their
raw device pointers are handed to ``libmate_b200.so``.  There is no CPU path.
"""

import ctypes

import numpy as np
import torch

from mate_b200 import _abi


class MateError(RuntimeError):
    pass


def _check(lib, rc):
    if rc != 0:
        msg = lib.mate_b200_last_error().decode('utf-8', 'replace')
        if rc == -1:
            raise ValueError(msg)
        raise MateError(f'mate_b200 error {rc}: {msg}')


def _dptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def rescale_tables(nc, nt, no):
    """(scale, shift) per observation column so that ``x * scale + shift`` equals the reference's
    ``normalize_observation`` (mate/agents/utils.py:97-127): subtract ``low`` where it is finite, then map the
    columns that are bounded on both sides (and not degenerate) to [-1, 1].  float32 [D][2] per team."""
    from mate_b200 import constants as consts  # pylint: disable=import-outside-toplevel

    tables = []
    for space in (consts.camera_observation_space_of(nc, nt, no), consts.target_observation_space_of(nc, nt, no)):
        low, high = np.asarray(space.low, dtype=np.float64), np.asarray(space.high, dtype=np.float64)
        below, above = np.isfinite(low), np.isfinite(high)
        both = below & above & (np.where(below & above, high - low, 0.0) > 0.0)
        span = np.where(both, high - low, 1.0)
        scale = np.where(both, 2.0 / span, 1.0)
        shift = np.where(below, -low, 0.0) * scale + np.where(both, -1.0, 0.0)
        tables.append(np.stack([scale, shift], axis=-1).astype(np.float32))
    return tuple(tables)


class BatchedSim:
    """One batch of MultiAgentTracking environments resident on one GPU."""

    def __init__(self, flat_config, num_envs, device=0, env_index_base=0):
        if not torch.cuda.is_available():
            raise RuntimeError('mate_b200 needs a CUDA device (there is no CPU fallback)')
        self.lib = _abi.load_library()
        self.cfg = dict(flat_config)
        self._cfg_struct = _abi.make_config_struct(self.cfg)
        self.B = int(num_envs)
        self.nc = int(self.cfg['num_cameras'])
        self.nt = int(self.cfg['num_targets'])
        self.no = int(self.cfg['num_obstacles'])
        index = device if isinstance(device, int) else (torch.device(device).index or 0)
        self.device = torch.device('cuda', index)
        self.handle = ctypes.c_void_p()
        _check(self.lib, self.lib.mate_b200_create(ctypes.byref(self._cfg_struct), self.B, self.device.index,
                                                   int(env_index_base), ctypes.byref(self.handle)))
        dc, dt = ctypes.c_int32(), ctypes.c_int32()
        _check(self.lib, self.lib.mate_b200_obs_dims(self.handle, ctypes.byref(dc), ctypes.byref(dt)))
        self.dc, self.dt = dc.value, dt.value
        dev = self.device
        self.cam_obs = torch.zeros((self.B, self.nc, self.dc), dtype=torch.float32, device=dev)
        self.tgt_obs = torch.zeros((self.B, self.nt, self.dt), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((self.B, 2), dtype=torch.float32, device=dev)
        self.done = torch.zeros((self.B,), dtype=torch.uint8, device=dev)
        self._stats = torch.zeros(16, dtype=torch.float32, device=dev)
        self._aux = None
        self._aux_struct = None
        self._zero_cam_act = torch.zeros((self.B, max(self.nc, 1), 2), dtype=torch.float32, device=dev)
        self._obs_ops, self._obs_ops_c, self._affine = [], None, None
        self._seed = 0

    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle:
            self.lib.mate_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def alloc_aux(self):
        """Device tensors for every MateStepAux field."""
        tmap = {np.uint8: torch.uint8, np.float32: torch.float32, np.int32: torch.int32}
        self._aux = {
            name: torch.zeros(shape(self.B, self.nc, self.nt, self.no), dtype=tmap[dtype], device=self.device)
            for name, _, dtype, shape in _abi.AUX_FIELDS
        }
        s = _abi.MateStepAux()
        for name, ctype, _, _ in _abi.AUX_FIELDS:
            setattr(s, name, ctypes.cast(ctypes.c_void_p(self._aux[name].data_ptr()), ctype))
        self._aux_struct = s
        return self._aux

    @property
    def aux(self):
        if self._aux is None:
            self.alloc_aux()
        return self._aux

    def _replay_struct(self, replay):
        if replay is None:
            return None, None
        transmit, choice = replay
        keep = []
        s = _abi.MateReplay()
        if transmit is not None:
            t = torch.as_tensor(np.ascontiguousarray(transmit, dtype=np.uint8)).to(self.device)
            t = t.reshape(self.B, self.nc, self.nt).contiguous()
            keep.append(t)
            s.transmit = ctypes.cast(ctypes.c_void_p(t.data_ptr()), _abi.c_uint8_p)
        if choice is not None:
            c = torch.as_tensor(np.ascontiguousarray(choice, dtype=np.int8)).to(self.device)
            c = c.reshape(self.B, self.nt).contiguous()
            keep.append(c)
            s.goal_choice = ctypes.cast(ctypes.c_void_p(c.data_ptr()), _abi.c_int8_p)
        return s, keep

    def _actions(self, cam_act, tgt_act):
        tgt_act = torch.as_tensor(tgt_act, dtype=torch.float32, device=self.device)
        tgt_act = tgt_act.reshape(self.B, self.nt, 2).contiguous()
        if self.nc:
            cam_act = torch.as_tensor(cam_act, dtype=torch.float32, device=self.device)
            cam_act = cam_act.reshape(self.B, self.nc, 2).contiguous()
        else:
            cam_act = self._zero_cam_act
        return cam_act, tgt_act

    # ------------------------------------------------------------------ observation wrappers (N1)
    def set_observation_wrappers(self, ops):
        """Observation wrapper codes (``_abi.OBS_*``), innermost first.  They are registered with the simulator
        (``mate_b200_set_observation_ops``): every reset / step / observe applies them to an environment's rows while
        the rows are still in shared memory, at no extra pass over the observation tensors."""
        ops = [int(op) for op in ops]
        if len(ops) > _abi.MAX_OBS_OPS:
            raise ValueError(f'at most {_abi.MAX_OBS_OPS} observation wrappers can be stacked')
        self._obs_ops = ops
        self._obs_ops_c = (ctypes.c_int32 * max(len(ops), 1))(*ops)
        if _abi.OBS_RESCALED in ops and self._affine is None:
            self._affine = tuple(torch.from_numpy(t).to(self.device) for t in rescale_tables(self.nc, self.nt, self.no))
        cam_aff, tgt_aff = self._affine if self._affine is not None else (None, None)
        _check(self.lib, self.lib.mate_b200_set_observation_ops(
            self.handle, self._obs_ops_c, len(ops), _dptr(cam_aff) if (self.nc and cam_aff is not None) else None, _dptr(tgt_aff)))

    def transform_observations(self):
        """The registered wrappers as a separate in-place pass over the current observation tensors
        (``mate_b200_transform_observations``); only for tensors that did NOT come out of reset / step / observe with
        the wrappers registered (those are transformed already)."""
        if not self._obs_ops:
            return
        cam_aff, tgt_aff = self._affine if self._affine is not None else (None, None)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_transform_observations(
                self.handle, _dptr(self.cam_obs), _dptr(self.tgt_obs), self._obs_ops_c, len(self._obs_ops),
                _dptr(cam_aff) if self.nc else None, _dptr(tgt_aff), self._stream()))

    def decode_actions(self, index, table):
        """DiscreteCamera / DiscreteTarget: int64 grid indices [B, N] -> float32 actions [B, N, 2]."""
        index = torch.as_tensor(index, device=self.device).to(torch.int64).contiguous()
        out = torch.empty(tuple(index.shape) + (2,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_decode_actions(_dptr(index), _dptr(table), int(table.shape[0]), _dptr(out),
                                                               index.numel(), self._stream()))
        return out

    # ------------------------------------------------------------------ API
    def seed(self, seed, env_mask=None):
        """``MultiAgentTracking.seed``: new key for the Philox streams, episode counters rewound (mate_b200_seed), so
        that the following resets are reproducible."""
        mask = None
        if env_mask is not None:
            mask = torch.as_tensor(env_mask, device=self.device).to(torch.uint8).contiguous()
        self._seed = int(seed)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_seed(self.handle, _dptr(mask), self._seed, self._stream()))

    def reset(self, seed=None, env_mask=None):
        """Reset the environments of ``env_mask`` (default: all).  With an explicit ``seed`` the reset is reproducible
        (``reset(seed=s)`` twice gives the same states, like the reference); without one it draws the next episodes
        of the current seed."""
        mask = None
        if env_mask is not None:
            mask = torch.as_tensor(env_mask, device=self.device).to(torch.uint8).contiguous()
        if seed is not None:
            self.seed(seed, mask)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_reset(self.handle, _dptr(mask), self._seed, _dptr(self.cam_obs),
                                                      _dptr(self.tgt_obs), self._stream()))
        return self.cam_obs, self.tgt_obs

    def step(self, cam_act, tgt_act, auto_reset=True, replay=None, aux=False):
        cam_act, tgt_act = self._actions(cam_act, tgt_act)
        rs, keep = self._replay_struct(replay)
        if aux and self._aux is None:
            self.alloc_aux()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_step(
                self.handle, _dptr(cam_act), _dptr(tgt_act), _dptr(self.cam_obs), _dptr(self.tgt_obs),
                _dptr(self.rewards), _dptr(self.done),
                ctypes.byref(self._aux_struct) if aux else None,
                ctypes.byref(rs) if rs is not None else None,
                _abi.MATE_STEP_AUTO_RESET if auto_reset else 0, self._stream()))
        del keep
        return (self.cam_obs, self.tgt_obs), self.rewards, self.done

    def observe(self, replay=None, aux=False):
        rs, keep = self._replay_struct(replay)
        if aux and self._aux is None:
            self.alloc_aux()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_observe(
                self.handle, _dptr(self.cam_obs), _dptr(self.tgt_obs),
                ctypes.byref(self._aux_struct) if aux else None,
                ctypes.byref(rs) if rs is not None else None, self._stream()))
        del keep
        return self.cam_obs, self.tgt_obs

    def fov_range(self, env, camera, angle_deg):
        """``Camera.sight_range_at`` (mate/entities.py:507-511) for equally shaped arrays of environment indices,
        camera indices and bearings in degrees; float64 tensor of the same shape."""
        env = torch.as_tensor(env, device=self.device).to(torch.int32).contiguous()
        camera = torch.as_tensor(camera, device=self.device).to(torch.int32).contiguous()
        angle = torch.as_tensor(angle_deg, device=self.device).to(torch.float64).contiguous()
        assert env.shape == camera.shape == angle.shape
        out = torch.empty_like(angle)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_fov_range(self.handle, _dptr(env), _dptr(camera), _dptr(angle), _dptr(out),
                                                          angle.numel(), self._stream()))
        return out

    def soft_coverage(self, auto_reset=True):
        """``AuxiliaryCameraRewards.compute_soft_coverage_scores`` for the current state and the camera-target masks
        of the last step: ``[B, Nc, Nt]`` float32 (mate_b200_soft_coverage).  With ``auto_reset`` the environments
        that were reset in the last step get zeros."""
        if self._aux is None:
            raise RuntimeError('soft_coverage needs a step / observe call with aux=True first')
        if getattr(self, '_soft', None) is None:
            self._soft = torch.zeros((self.B, self.nc, self.nt), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_soft_coverage(
                self.handle, _dptr(self._aux['mask_ct']), _dptr(self.done) if auto_reset else None, _dptr(self._soft), self._stream()))
        return self._soft

    def auxiliary_terms(self, soft_matrix=None):
        """Per-agent terms of the auxiliary-reward / training-information wrappers for the LAST step
        (mate_b200_auxiliary_terms): ``(cam_terms [B, Nc, CAM_TERMS], tgt_terms [B, Nt, TGT_TERMS])``; the
        ``soft_coverage_score`` columns are filled when ``soft_matrix`` (from :meth:`soft_coverage`) is given."""
        if self._aux is None:
            raise RuntimeError('auxiliary_terms needs a step / observe call with aux=True first')
        if getattr(self, '_terms', None) is None:
            self._terms = (torch.zeros((self.B, self.nc, _abi.CAM_TERMS), dtype=torch.float32, device=self.device),
                           torch.zeros((self.B, self.nt, _abi.TGT_TERMS), dtype=torch.float32, device=self.device))
        cam_terms, tgt_terms = self._terms
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_auxiliary_terms(
                self.handle, ctypes.byref(self._aux_struct), _dptr(self.rewards), _dptr(soft_matrix),
                _dptr(cam_terms) if self.nc else None, _dptr(tgt_terms), self._stream()))
        return cam_terms, tgt_terms

    def step_host(self, cam_act, tgt_act, out, auto_reset=True, rows_kept=False):
        """Host-buffer step: `cam_act`/`tgt_act` and the tensors in `out` (cam_obs, tgt_obs,
        rewards, done) are CPU tensors (pinned for full PCIe speed).  `rows_kept=True` promises that
        `out` is the tuple of the previous `step_host` call and has not been written to since
        (MATE_STEP_HOST_ROWS_KEPT, include/mate_b200.h): parts of the rows that stay zero are then not rewritten."""
        cam_obs, tgt_obs, rewards, done = out
        flags = (_abi.MATE_STEP_AUTO_RESET if auto_reset else 0) | (_abi.MATE_STEP_HOST_ROWS_KEPT if rows_kept else 0)
        _check(self.lib, self.lib.mate_b200_step_host(
            self.handle, _dptr(cam_act) if self.nc else None, _dptr(tgt_act),
            _dptr(cam_obs) if self.nc else None, _dptr(tgt_obs), _dptr(rewards), _dptr(done), flags))
        return out

    def get_state(self):
        arrays = _abi.alloc_state_arrays(self.B, self.nc, self.nt, self.no)
        view = _abi.state_view_from_arrays(arrays)
        _check(self.lib, self.lib.mate_b200_get_state(self.handle, ctypes.byref(view)))
        return arrays

    def set_state(self, arrays):
        arrays = dict(arrays)
        view = _abi.state_view_from_arrays(arrays)
        _check(self.lib, self.lib.mate_b200_set_state(self.handle, ctypes.byref(view)))

    def episode_stats(self, reset_after=False):
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.mate_b200_episode_stats(self.handle, _dptr(self._stats), int(reset_after),
                                                              self._stream()))
        return self._stats

    def host_leg_info(self):
        """(leg, row bytes on the link) of the last `step_host` call: 0 dense copy, 1 rows rebuilt from their non-zero
        chunks, 2 changes patched into kept rows (mate_b200_host_leg_info)."""
        leg, link = ctypes.c_int32(0), ctypes.c_uint64(0)
        _check(self.lib, self.lib.mate_b200_host_leg_info(self.handle, ctypes.byref(leg), ctypes.byref(link)))
        return int(leg.value), int(link.value)

    @property
    def launch_count(self):
        return int(self.lib.mate_b200_launch_count(self.handle))
