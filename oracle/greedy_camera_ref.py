"""NumPy restatement of ``GreedyCameraAgent`` (mate/agents/greedy.py:14-232) for a whole team, driven like
``group_step`` (mate/wrappers/single_team.py:79-92: observe -> communicate -> act) -- TEST INFRASTRUCTURE: the
checker of the CUDA kernel, itself pinned against recorded reference agents in ``tests/golden/agents_*.npz``."""

import numpy as np

MEMORY_PERIOD = 25
RANGE_FACTOR = 1.1


def norm_angle(a):
    return (a + 180.0) % 360.0 - 180.0


def act_from_target(cam_xy, phi, theta, target_xy, min_view, rmax, rot_step, zoom_step):
    """act_from_target_states for the selected target (greedy.py:117-154)."""
    rel = target_xy - cam_xy
    distance = float(np.hypot(rel[0], rel[1]))
    best_orientation = np.rad2deg(np.arctan2(rel[1], rel[0]))
    area_product = min_view * rmax * rmax
    if distance * (1.0 + np.sin(np.deg2rad(min_view / 2.0))) >= rmax:
        best_view = min_view
    elif distance <= np.sqrt(area_product / 180.0) / 2.0:
        best_view = 180.0
    else:
        best = 180.0
        for _ in range(20):
            sight_range = distance * (1.0 + np.sin(np.deg2rad(min(best / 2.0, 90.0))))
            best = area_product / (sight_range * sight_range)
        best_view = float(np.clip(best, min_view, 180.0))
    action = np.array([norm_angle(best_orientation - phi), best_view - theta])
    return np.clip(action, [-rot_step, -zoom_step], [rot_step, zoom_step])


def team_step(cam_xy, cam_phi, cam_theta, tgt_state, tracked, mem, draws, cfg):
    """One group_step of the Nc camera agents of one environment.

    tgt_state [Nt, 4] public target states (x, y, sight range, is_loaded), tracked [Nc, Nt] bool (the flags of the
    cameras' observations); mem: dict of arrays BEFORE the step (memory [Nc, Nt, 4], time2forget [Nc, Nt],
    never_loaded [Nc, Nt], prev_action [Nc, 2], delay [Nc, Nc], neighbors [Nc] bit sets, has_state [Nc]);
    draws: dict(binomial [Nc], sample [Nc, 2], delay [Nc, Nc]) recorded outcomes; cfg: (min_view, rmax, rot, zoom).
    Returns (actions [Nc, 2], memory after)."""
    min_view, rmax, rot_step, zoom_step = cfg
    nc, nt = tracked.shape
    memory, t2f, never = mem['memory'].copy(), mem['time2forget'].copy(), mem['never_loaded'].copy()
    delay, neighbors, has_state = mem['delay'].copy(), mem['neighbors'].copy(), mem['has_state'].copy()
    # observe -> process_messages (greedy.py:104-115)
    outbox = [[] for _ in range(nc)]
    for c in range(nc):
        t2f[c] = np.maximum(t2f[c] - 1, 0)
        for t in np.flatnonzero(tracked[c]):
            t2f[c, t] = MEMORY_PERIOD
            memory[c, t] = tgt_state[t]
            if tgt_state[t, 3] != 0:
                never[c, t] = 0
            outbox[c].append(t)
    # send_responses (greedy.py:156-194); all agents send before anyone receives (single_team.py:55-59)
    messages = []   # (sender, recipient, has_state, [target indices])
    for c in range(nc):
        delay[c] = np.maximum(delay[c] - 1, 0)
        if has_state[c] or outbox[c]:
            for k in range(nc):
                if k == c or delay[c, k] > 0:
                    continue
                targets = []
                if outbox[c] and (neighbors[c] >> k) & 1:
                    threshold = RANGE_FACTOR * rmax
                    targets = [t for t in outbox[c] if np.hypot(*(tgt_state[t, :2] - cam_xy[k])) < threshold]
                if has_state[c] or targets:
                    messages.append((c, k, bool(has_state[c]), targets))
                    assert draws['delay'][c, k] >= 0
                    delay[c, k] = draws['delay'][c, k]
            has_state[c] = 0
    # receive_responses (greedy.py:196-232)
    for sender, recipient, with_state, targets in messages:
        if with_state:
            neighbors[recipient] |= 1 << sender    # greedy.py:219 adds the sender unconditionally
        for t in targets:
            memory[recipient, t] = tgt_state[t]
            t2f[recipient, t] = MEMORY_PERIOD
            if tgt_state[t, 3] != 0:
                never[recipient, t] = 0
    # act (greedy.py:68-102)
    actions = np.zeros((nc, 2))
    for c in range(nc):
        candidates = [t for t in np.flatnonzero(t2f[c]) if np.hypot(*(memory[c, t, :2] - cam_xy[c])) < RANGE_FACTOR * rmax]
        if candidates:
            nearest = min(candidates, key=lambda t: np.hypot(*(memory[c, t, :2] - cam_xy[c])))
            actions[c] = act_from_target(cam_xy[c], cam_phi[c], cam_theta[c], memory[c, nearest, :2], min_view, rmax, rot_step, zoom_step)
        elif draws['binomial'][c] == 1:
            actions[c] = draws['sample'][c]
        else:
            actions[c] = mem['prev_action'][c]
    return actions, {'memory': memory, 'time2forget': t2f, 'never_loaded': never, 'prev_action': actions.copy(), 'delay': delay,
                     'neighbors': neighbors, 'has_state': has_state}
