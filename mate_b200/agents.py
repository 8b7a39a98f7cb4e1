"""Batched rule-based opponents.

``GreedyTargetAgent`` mirrors the reference's constructor (mate/agents/greedy.py:241-256).  One instance stands for
the whole target team of every environment of a batch: its memory (goal, remembered non-empty warehouses, previous
location, previous noise per target) is a ``[B, Nt, 6]`` CUDA tensor and one kernel per step
(``mate_b200_greedy_target_actions``) runs observe -> communicate -> act for all of them.  The agents' random draws
are counter-based (Philox), keyed on ``seed``, the global environment index and the step.
"""

import ctypes

import numpy as np
import torch

from mate_b200 import _abi
from mate_b200.sim import _check, _dptr


class TargetAgentBase:   # marker base, like mate.agents.base.TargetAgentBase
    TEAM = 'target'


class CameraAgentBase:   # marker base, like mate.agents.base.CameraAgentBase
    TEAM = 'camera'


class GreedyTargetAgent(TargetAgentBase):
    """Greedy Target Agent: runs towards the destination (desired warehouse) with some noise."""

    def __init__(self, seed=None, noise_scale=0.5):
        self.noise_scale = float(noise_scale)
        self._seed = 0
        self.seed(seed)
        self.memory = None
        self.actions = None
        self._serial = 0
        self._sim = None

    def seed(self, seed=None):
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 63))
        self._seed = int(seed)
        return [self._seed]

    def clone(self):
        return GreedyTargetAgent(seed=self._seed + 1, noise_scale=self.noise_scale)

    def spawn(self, num_agents):   # one batched instance plays every agent of the team
        return [self] * num_agents

    def bind(self, sim):
        """Allocate the team memory for a simulator (``mate_b200.sim.BatchedSim``)."""
        self._sim = sim
        # stored field-major [6, Nt, B] (coalesced in the kernel, one thread per environment); `memory` is its [B, Nt, 6] view
        self._memory = torch.zeros((_abi.AGENT_MEMORY, sim.nt, sim.B), dtype=torch.float64, device=sim.device)
        self.memory = self._memory.permute(2, 1, 0)
        self.actions = torch.zeros((sim.B, sim.nt, 2), dtype=torch.float32, device=sim.device)
        self._serial = 0

    def act(self, reset_mask=None, replay=None):
        """Joint target action ``[B, Nt, 2]`` for the simulator's current state.  ``reset_mask`` ``[B]`` (uint8 / bool
        CUDA tensor, or ``True`` for all): environments whose agents are reset first.  ``replay``: dict with the
        recorded draws ``binomial`` / ``sample`` / ``choice`` / ``reset_sample`` (parity tests)."""
        sim = self._sim
        if reset_mask is True:
            reset_mask = torch.ones(sim.B, dtype=torch.uint8, device=sim.device)
        elif reset_mask is not None:
            reset_mask = torch.as_tensor(reset_mask, device=sim.device).to(torch.uint8).contiguous()
        rs, keep = None, []
        if replay is not None:
            rs = _abi.MateAgentReplay()
            for name, dtype, ctype in (('binomial', torch.uint8, _abi.c_uint8_p), ('sample', torch.float64, _abi.c_double_p),
                                       ('choice', torch.int8, _abi.c_int8_p), ('reset_sample', torch.float64, _abi.c_double_p)):
                if replay.get(name) is not None:
                    t = torch.as_tensor(np.ascontiguousarray(replay[name])).to(dtype).to(sim.device).contiguous()
                    keep.append(t)
                    setattr(rs, name, ctypes.cast(ctypes.c_void_p(t.data_ptr()), ctype))
        with torch.cuda.device(sim.device):
            _check(sim.lib, sim.lib.mate_b200_greedy_target_actions(
                sim.handle, _dptr(self._memory), _dptr(reset_mask), self.noise_scale, self._seed % (2 ** 64), self._serial,
                ctypes.byref(rs) if rs is not None else None, _dptr(self.actions), sim._stream()))  # pylint: disable=protected-access
        del keep
        self._serial += 1
        return self.actions


class GreedyCameraAgent(CameraAgentBase):
    """Greedy Camera Agent (mate/agents/greedy.py:14-232): tracks the nearest remembered target; without one it keeps
    its previous action or, with probability 0.1, draws a new random one.  One instance stands for the camera team of
    every environment of a batch (``mate_b200_greedy_camera_actions``).  Only the reference's default arguments are
    supported (``memory_period=25``, ``filterout_unloaded=False``, ``filterout_beyond_range=True``)."""

    def __init__(self, seed=None, memory_period=25, filterout_unloaded=False, filterout_beyond_range=True):
        if memory_period != 25 or filterout_unloaded or not filterout_beyond_range:
            raise NotImplementedError('the batched GreedyCameraAgent supports the default arguments of the reference only')
        self._seed = 0
        self.seed(seed)
        self.memory = None
        self.actions = None
        self._serial = 0
        self._sim = None

    def seed(self, seed=None):
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 63))
        self._seed = int(seed)
        return [self._seed]

    def clone(self):
        return GreedyCameraAgent(seed=self._seed + 1)

    def spawn(self, num_agents):
        return [self] * num_agents

    def bind(self, sim):
        self._sim = sim
        width = 6 * sim.nt + sim.nc + 4
        # stored field-major [M, B, Nc] (the kernel's accesses are coalesced); `memory` is the [B, Nc, M] view of it
        self._memory = torch.zeros((width, sim.B, sim.nc), dtype=torch.float64, device=sim.device)
        self.memory = self._memory.permute(1, 2, 0)
        self.actions = torch.zeros((sim.B, sim.nc, 2), dtype=torch.float32, device=sim.device)
        self._serial = 0

    def act(self, tracked, reset_mask=None, replay=None):
        """Joint camera action ``[B, Nc, 2]``.  ``tracked`` ``[B, Nc, Nt]``: the target flags of the cameras' current
        observations; ``reset_mask`` / ``replay`` as for :class:`GreedyTargetAgent` (draws: ``binomial`` int8 with -1
        for "not drawn", ``sample``, ``delay`` int32 ``[B, Nc, Nc]``)."""
        sim = self._sim
        tracked = torch.as_tensor(tracked, device=sim.device).to(torch.uint8).contiguous()
        if reset_mask is True:
            reset_mask = torch.ones(sim.B, dtype=torch.uint8, device=sim.device)
        elif reset_mask is not None:
            reset_mask = torch.as_tensor(reset_mask, device=sim.device).to(torch.uint8).contiguous()
        rs, keep = None, []
        if replay is not None:
            rs = _abi.MateCameraAgentReplay()
            for name, dtype, ctype in (('binomial', torch.int8, _abi.c_int8_p), ('sample', torch.float64, _abi.c_double_p),
                                       ('delay', torch.int32, _abi.c_int32_p)):
                if replay.get(name) is not None:
                    t = torch.as_tensor(np.ascontiguousarray(replay[name])).to(dtype).to(sim.device).contiguous()
                    keep.append(t)
                    setattr(rs, name, ctypes.cast(ctypes.c_void_p(t.data_ptr()), ctype))
        with torch.cuda.device(sim.device):
            _check(sim.lib, sim.lib.mate_b200_greedy_camera_actions(
                sim.handle, _dptr(self._memory), _dptr(tracked), _dptr(reset_mask), self._seed % (2 ** 64), self._serial,
                ctypes.byref(rs) if rs is not None else None, _dptr(self.actions), sim._stream()))  # pylint: disable=protected-access
        del keep
        self._serial += 1
        return self.actions
