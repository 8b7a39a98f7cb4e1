"""How often does the reference's rounding-noise decision at an exactly tangent field-of-view ray (DESIGN.md "Tangent
rays") change a camera -> target mask bit?  Test infrastructure, build container only (runs the UNMODIFIED reference
through oracle/gymshim like gen_golden.py).

Records fresh traces of the reference (other seeds than the committed fixtures, not kept), replays every step through
the C oracle from the reference's own state and recorded draws, and counts the steps whose `camera_target_view_mask`
differs -- checking that every differing (camera, target) pair lies in a tangent sliver (tests/golden_util.py).

    python oracle/measure_tangent_rate.py --steps 10050 --seeds 20 21 22 23 --config MATE-4v8-9.yaml --out profiles/r2u_tangent_rate.txt
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)                          # gen_golden
sys.path.insert(0, os.path.join(REPO, 'tests'))   # golden_util
sys.path.insert(0, REPO)                          # the `oracle` package itself must win over oracle/oracle.py


def main():
    import gen_golden  # pylint: disable=import-outside-toplevel
    import golden_util as gu  # pylint: disable=import-outside-toplevel
    from oracle.oracle import Oracle  # pylint: disable=import-outside-toplevel

    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument('--config', nargs='+', default=['MATE-4v8-9.yaml'])
    parser.add_argument('--seeds', nargs='+', type=int, default=[20, 21])
    parser.add_argument('--steps', type=int, default=10050)
    parser.add_argument('--policy', default='random')
    parser.add_argument('--out', default=None)
    args = parser.parse_args()
    mate = gen_golden._import_reference()  # pylint: disable=protected-access
    lines = []
    total_steps = total_flip_steps = total_pairs = 0
    for config in args.config:
        for seed in args.seeds:
            t0 = time.time()
            with tempfile.TemporaryDirectory() as tmp:
                path = os.path.join(tmp, 'trace.npz')
                gen_golden.run_trace(mate, config, seed, args.steps, args.policy, 1 << 30, path)
                g = np.load(path)
                cfg = gu.flat_config(g)
                sim = Oracle(cfg, 1)
                sim.set_state(gu.state_arrays(g))
                aux = sim.alloc_aux()
                T = int(g['num_steps'])
                flip_steps, pairs, other = [], 0, 0
                for k in range(T):
                    sim.step(g['step_cam_act'][k][None], g['step_tgt_act'][k][None], transmit=g['step_transmit'][k][None],
                             goal_choice=g['step_goal_choice'][k][None], aux=aux)
                    diff = np.argwhere(aux['mask_ct'][0] != g['step_mask_ct'][k])
                    if len(diff):
                        flip_steps.append(k)
                        for c, t in diff:
                            pairs += 1
                            if not gu.in_tangent_sliver(g['init_cam_xy'][c], g['step_tgt_xy'][k][t], g['init_obs_xyr'],
                                                        cfg['camera_max_sight_range']):
                                other += 1
                    for key in ('mask_cc', 'mask_tc', 'mask_to', 'mask_tt'):
                        if not (aux[key][0] == g['step_' + key][k]).all():
                            other += 1
                    # the next step starts from the reference's state (also where nothing differed: no drift)
                    sim.set_state(gu.step_state_arrays(g, k))
            line = (f'{config} seed {seed} policy {args.policy}: {T} steps, {len(flip_steps)} steps with a differing camera->target bit '
                    f'({pairs} pairs, all in a tangent sliver: {other == 0}) at steps {flip_steps[:12]}  [{time.time() - t0:.0f} s]')
            print(line, flush=True)
            lines.append(line)
            total_steps += T
            total_flip_steps += len(flip_steps)
            total_pairs += pairs
            assert other == 0, line
    summary = (f'TOTAL: {total_flip_steps} of {total_steps} reference steps differ in camera_target_view_mask '
               f'({total_pairs} (camera, target) pairs), every one inside a tangent sliver; all other masks equal on every step')
    print(summary)
    if args.out:
        with open(args.out, 'a', encoding='utf-8') as f:
            f.write('\n'.join(lines + [summary]) + '\n')


if __name__ == '__main__':
    main()
