"""make_target kernel vs the oracle (itself pinned bit-exact to the reference's ReplayBuffer.make_target)."""
import numpy as np
import pytest
import torch

import _tables as T
from oracle import rng, targets as otargets

pytestmark = pytest.mark.gpu


class _Cfg:
    pass


def test_make_target_bit_exact():
    from muzero_hypermodel_b200.replay_buffer import DeviceGames, make_target_batch
    z = T.load("targets")
    rs = np.random.RandomState(0)
    groups = {}
    for i in range(int(z["n"])):
        groups.setdefault(tuple(z[f"{i}/cfg"]), []).append(i)
    checked = 0
    for (K, td, disc, A, players), idx in groups.items():
        K, td, A = int(K), int(td), int(A)
        disc = int(disc) if float(disc).is_integer() else float(disc)
        cfg = _Cfg(); cfg.num_unroll_steps, cfg.td_steps, cfg.discount = K, td, disc
        # entry arrays of all games of this group; visit COUNTS (the device stores counts, policy = count / sum)
        rew, tp, rv, vis, act, start, length, rean = [], [], [], [], [], [], [], []
        use_rean = all(len(z[f"{i}/reanalysed"]) for i in idx[:1])
        exp = []
        for i in idx:
            pre = f"{i}/"
            n = len(z[pre + "root_values"])
            counts = rs.multinomial(50, z[pre + "child_visits"][0] * 0 + 1.0 / A, size=n).astype(np.int64)
            r32 = z[pre + "reward_history"].astype(np.float32)
            start.append(sum(len(x) for x in rew)); length.append(n)
            rew.append(r32); tp.append(z[pre + "to_play_history"]); act.append(z[pre + "action_history"])
            rv.append(np.concatenate([z[pre + "root_values"], [0.0]]))
            re = z[pre + "reanalysed"]
            rean.append(np.concatenate([re if len(re) else z[pre + "root_values"], [0.0]]))
            vis.append(np.concatenate([counts, np.zeros((1, A), dtype=np.int64)]))
            slot, step = [int(v) for v in z[pre + "slot_step"]]
            past = list(range(1000))
            e = otargets.make_target(z[pre + "root_values"].tolist(), [float(x) for x in r32], z[pre + "to_play_history"].tolist(),
                                     [[int(c) / int(row.sum()) if c else 0 for c in row] for row in counts],
                                     z[pre + "action_history"].tolist(), int(z[pre + "state_index"]), K, td, disc, A,
                                     reanalysed_root_values=re.tolist() if len(re) else None,
                                     pad_action=lambda row: rng.pad_action(T.SEED, slot, step, past.pop(0), A))
            exp.append((e, slot, step, int(z[pre + "state_index"]), len(re) > 0))
        for with_rean in (False, True):
            sel = [k for k, e in enumerate(exp) if e[4] == with_rean]
            if not sel:
                continue
            games = DeviceGames(np.concatenate(rew), np.concatenate(tp), np.concatenate(rv), np.concatenate(vis),
                                np.concatenate(act), start, length, reanalysed=np.concatenate(rean) if with_rean else None)
            tv, tr, tpol, ta = make_target_batch(games, sel, [exp[k][3] for k in sel], cfg, seed=T.SEED,
                                                 batch_slot=[exp[k][1] for k in sel], batch_step=[exp[k][2] for k in sel])
            tv, tr, tpol, ta = tv.cpu().numpy(), tr.cpu().numpy(), tpol.cpu().numpy(), ta.cpu().numpy()
            for j, k in enumerate(sel):
                ev, er, ep, ea = exp[k][0]
                assert np.array(ev, dtype=np.float64).tobytes() == tv[j].tobytes(), (k, "value")
                assert np.array(er, dtype=np.float64).tobytes() == tr[j].tobytes(), (k, "reward")
                assert np.array(ep, dtype=np.float64).tobytes() == tpol[j].tobytes(), (k, "policy")
                np.testing.assert_array_equal(np.array(ea, dtype=np.int32), ta[j])
                checked += 1
    assert checked == int(z["n"])
