"""CPU restatement of the replay-buffer target builder (test infrastructure).

Follows /root/reference/replay_buffer.py: compute_target_value :222-254, make_target :256-295.
Works on plain lists (the GameHistory fields of self_play.py:485-495) and keeps the reference's
float64 left-to-right accumulation order, including its indexing quirk: the sign of reward
`reward_history[index+1+i]` is decided by `to_play_history[index+i]` (:246-250).
Pinned by tests/golden/targets_*.npz (reference ReplayBuffer.make_target outputs).
"""


def compute_target_value(root_values, reward_history, to_play_history, index, td_steps, discount,
                         reanalysed_root_values=None):
    bootstrap = index + td_steps
    if bootstrap < len(root_values):
        rv = root_values if reanalysed_root_values is None else reanalysed_root_values
        last = rv[bootstrap] if to_play_history[bootstrap] == to_play_history[index] else -rv[bootstrap]
        value = last * discount ** td_steps
    else:
        value = 0
    for i, reward in enumerate(reward_history[index + 1:bootstrap + 1]):
        signed = reward if to_play_history[index] == to_play_history[index + i] else -reward
        value += signed * discount ** i
    return value


def make_target(root_values, reward_history, to_play_history, child_visits, action_history,
                state_index, num_unroll_steps, td_steps, discount, n_actions,
                reanalysed_root_values=None, pad_action=lambda row: 0):
    """Returns (target_values, target_rewards, target_policies, actions), one row per unroll step.

    pad_action(row) injects the reference's `numpy.random.choice(action_space)` for rows past the
    end of the game (:291)."""
    values, rewards, policies, actions = [], [], [], []
    n = len(root_values)
    width = len(child_visits[0])
    for row, cur in enumerate(range(state_index, state_index + num_unroll_steps + 1)):
        v = compute_target_value(root_values, reward_history, to_play_history, cur, td_steps, discount,
                                 reanalysed_root_values)
        if cur < n:
            values.append(v)
            rewards.append(reward_history[cur])
            policies.append(list(child_visits[cur]))
            actions.append(action_history[cur])
        elif cur == n:
            values.append(0)
            rewards.append(reward_history[cur])
            policies.append([1 / width] * width)
            actions.append(action_history[cur])
        else:
            values.append(0)
            rewards.append(0)
            policies.append([1 / width] * width)
            actions.append(pad_action(row))
    return values, rewards, policies, actions
