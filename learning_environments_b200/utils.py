"""Host-side helpers with the reference's names and behaviour (/root/reference/utils.py: ReplayBuffer :9-72,
AverageMeter :75-105, one-hot helpers :108-130, save_lists :144-160), built for this package's data layout.

In the fused path the replay ring lives in HBM inside the kernel workspace as packed rows `[s | a | s' | r | d]`; the
ReplayBuffer below keeps the same packed-row layout on the host (one tensor, the five reference attributes are column
views of it), so a ring copied back from the device is a ReplayBuffer without re-packing, and learn() of the
step-by-step API hands rows to le_td_update unchanged.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F


class ReplayBuffer(object):
    """Ring of transitions.  Public surface of the reference class: add / sample / get_all / merge_buffer / merge_vectors /
    get_size / clear and the attributes state, action, next_state, reward, done, ptr, size, max_size."""

    _INITIAL_ROWS = 4096      # the reference zero-fills max_size rows up front (44 MB per test() call at 1e6 rows); this grows on demand

    def __init__(self, state_dim, action_dim, device, max_size=int(1e6)):
        self.device, self.max_size = device, int(max_size)
        self.state_dim, self.action_dim = state_dim, action_dim
        self.ptr = self.size = 0
        self._alloc(min(self.max_size, self._INITIAL_ROWS))

    # ---- storage: one packed tensor, five column views --------------------------------------------------------
    def _alloc(self, rows):
        sd, ad = self.state_dim, self.action_dim
        self._rows = int(rows)
        self._packed = torch.zeros((self._rows, 2 * sd + ad + 2))
        cuts = np.cumsum([0, sd, ad, sd, 1, 1])
        self.state, self.action, self.next_state, self.reward, self.done = (self._packed[:, a:b] for a, b in zip(cuts[:-1], cuts[1:]))

    def _ensure_row(self, row):
        if row < self._rows:
            return
        keep = self._packed
        self._alloc(min(self.max_size, max(4 * self._rows, row + 1)))
        self._packed[:keep.shape[0]] = keep

    def _fields(self):
        return (self.state, self.action, self.next_state, self.reward, self.done)

    # ---- reference API ----------------------------------------------------------------------------------------------
    def add(self, state, action, next_state, reward, done):
        self._ensure_row(self.ptr)
        for column, value in zip(self._fields(), (state, action, next_state, reward, done)):
            column[self.ptr] = torch.as_tensor(value).detach().reshape(-1)
        self.ptr = (self.ptr + 1) % self.max_size          # utils.py:31-32
        self.size = min(self.size + 1, self.max_size)

    def sample(self, batch_size):
        return self._sample_idx(np.random.randint(0, self.size, size=batch_size))     # with replacement (utils.py:35)

    def _sample_idx(self, idx):
        return tuple(column[idx].to(self.device).detach() for column in self._fields())

    def get_all(self):
        return tuple(column[:self.size].to(self.device).detach() for column in self._fields())

    def merge_buffer(self, other_replay_buffer):
        self.merge_vectors(*other_replay_buffer.get_all())

    def merge_vectors(self, states, actions, next_states, rewards, dones):
        for transition in zip(states, actions, next_states, rewards, dones):
            self.add(*transition)

    def get_size(self):
        return self.size

    def clear(self):
        self.ptr = self.size = 0
        self._alloc(min(self.max_size, self._INITIAL_ROWS))


class AverageMeter(object):
    """Running list of values with windowed means (utils.py:75-105); `vals` is the raw list train()/test() return."""

    def __init__(self, print_str):
        self.print_str, self.vals, self.it = print_str, [], 0

    def _window(self, num, skip):
        hi = max(len(self.vals) - skip, 0)
        return self.vals[max(hi - num, 0):hi]

    def _mean(self, num, ignore_last):
        window = self._window(num, ignore_last)
        return sum(window) / (len(window) + 1e-9)          # the reference's guarded division (utils.py:105)

    def update(self, val, print_rate=10):
        self.vals.append(val.item() if torch.is_tensor(val) else val)
        self.it += 1
        if self.it % print_rate == 0:
            print(self.print_str + "{:15.6f} {:>25} {}".format(self._mean(print_rate, 0), "Total updates: ", self.it))

    def get_mean(self, num=10):
        return self._mean(num, 0)

    def get_mean_last(self, num=10):
        return self._mean(num, num)

    def get_raw_data(self):
        return self.vals


def to_one_hot_encoding(normal, one_hot_dim):
    """Scalar (python number or 0-d / 1-element tensor) -> [one_hot_dim]; 1-D index vector -> [n, one_hot_dim]; float32."""
    index = normal.squeeze() if torch.is_tensor(normal) else torch.as_tensor(int(normal))
    if index.dim() > 1:
        raise NotImplementedError('One hot encoding supported only for scalar values and 1D vectors')
    return F.one_hot(index.long(), int(one_hot_dim)).to(torch.float32)


def from_one_hot_encoding(one_hot):
    return torch.argmax(one_hot).reshape(1)


def save_lists(mode, config, reward_list, train_steps_needed, episode_length_needed, env_reward_overview, experiment_name=None,
               out_dir=None):
    """Result file of the vary_hp evaluators (utils.py:144-160): `<mode>_<experiment_name>.pt` with the same dict keys, so the
    reference's plot scripts load it.  Returns the path."""
    import pandas as pd
    path = os.path.join(out_dir or os.getcwd(), "%s_%s.pt" % (mode, experiment_name or "_experiment_"))
    torch.save(dict(config=config, reward_list=reward_list, train_steps_needed=train_steps_needed,
                    episode_length_needed=episode_length_needed,
                    env_reward_overview=pd.DataFrame.from_dict(env_reward_overview, orient="index")), path)
    return path
