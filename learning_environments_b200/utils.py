"""Host-side helpers that keep the reference's names and behaviour (/root/reference/utils.py):
ReplayBuffer (:9-72), AverageMeter (:75-105), one-hot helpers (:108-130), save_lists (:144-160).

In the fused path the replay ring lives in HBM inside the kernel's workspace; this ReplayBuffer is the
drop-in object that agent.train()/test() return and that learn() samples from in the step-by-step API.
"""
import os

import numpy as np
import torch


class ReplayBuffer(object):
    def __init__(self, state_dim, action_dim, device, max_size=int(1e6)):
        self.device = device
        self.max_size = int(max_size)
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.ptr = 0
        self.size = 0
        self._alloc(min(self.max_size, 4096))

    def _alloc(self, rows):
        # grown on demand: the reference zero-fills max_size rows up front (44 MB per test() call at 1e6 rows)
        self._rows = rows
        self.state = torch.zeros((rows, self.state_dim))
        self.action = torch.zeros((rows, self.action_dim))
        self.next_state = torch.zeros((rows, self.state_dim))
        self.reward = torch.zeros((rows, 1))
        self.done = torch.zeros((rows, 1))

    def _grow(self):
        rows = min(self.max_size, self._rows * 4)
        old = (self.state, self.action, self.next_state, self.reward, self.done)
        n = self._rows
        self._alloc(rows)
        for dst, src in zip((self.state, self.action, self.next_state, self.reward, self.done), old):
            dst[:n] = src

    def add(self, state, action, next_state, reward, done):
        if self.ptr >= self._rows:
            self._grow()
        self.state[self.ptr] = torch.as_tensor(state).detach().reshape(-1)
        self.action[self.ptr] = torch.as_tensor(action).detach().reshape(-1)
        self.next_state[self.ptr] = torch.as_tensor(next_state).detach().reshape(-1)
        self.reward[self.ptr] = torch.as_tensor(reward).detach().reshape(-1)
        self.done[self.ptr] = torch.as_tensor(done).detach().reshape(-1)
        self.ptr = (self.ptr + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def sample(self, batch_size):
        idx = np.random.randint(0, self.size, size=batch_size)   # with replacement, as the reference
        return self._sample_idx(idx)

    def _sample_idx(self, idx):
        return tuple(t[idx].to(self.device).detach() for t in (self.state, self.action, self.next_state, self.reward, self.done))

    def get_all(self):
        return tuple(t[:self.size].to(self.device).detach() for t in (self.state, self.action, self.next_state, self.reward, self.done))

    def merge_buffer(self, other_replay_buffer):
        states, actions, next_states, rewards, dones = other_replay_buffer.get_all()
        self.merge_vectors(states=states, actions=actions, next_states=next_states, rewards=rewards, dones=dones)

    def merge_vectors(self, states, actions, next_states, rewards, dones):
        for i in range(len(states)):
            self.add(states[i], actions[i], next_states[i], rewards[i], dones[i])

    def get_size(self):
        return self.size

    def clear(self):
        self.__init__(state_dim=self.state_dim, action_dim=self.action_dim, device=self.device, max_size=self.max_size)


class AverageMeter(object):
    def __init__(self, print_str):
        self.print_str = print_str
        self.vals = []
        self.it = 0

    def update(self, val, print_rate=10):
        if torch.is_tensor(val):
            val = val.item()
        self.vals.append(val)
        self.it += 1
        if self.it % print_rate == 0:
            mean_val = self._mean(num=print_rate, ignore_last=0)
            print(self.print_str + "{:15.6f} {:>25} {}".format(mean_val, "Total updates: ", self.it))

    def get_mean(self, num=10):
        return self._mean(num, ignore_last=0)

    def get_mean_last(self, num=10):
        return self._mean(num, ignore_last=num)

    def get_raw_data(self):
        return self.vals

    def _mean(self, num, ignore_last):
        vals = self.vals[max(len(self.vals) - num - ignore_last, 0): max(len(self.vals) - ignore_last, 0)]
        return sum(vals) / (len(vals) + 1e-9)


def to_one_hot_encoding(normal, one_hot_dim):
    if torch.is_tensor(normal):
        normal = normal.squeeze()
    if not torch.is_tensor(normal):
        one_hot = torch.zeros(one_hot_dim)
        one_hot[int(normal)] = 1
    elif normal.dim() == 0 or (normal.dim() == 1 and len(normal) == 1):
        one_hot = torch.zeros(one_hot_dim)
        one_hot[int(normal.item())] = 1
    elif normal.dim() == 1:
        one_hot = torch.zeros(len(normal), one_hot_dim)
        one_hot[torch.arange(len(normal)), normal.long()] = 1
    else:
        raise NotImplementedError('One hot encoding supported only for scalar values and 1D vectors')
    return one_hot


def from_one_hot_encoding(one_hot):
    return torch.tensor([torch.argmax(one_hot)])


def save_lists(mode, config, reward_list, train_steps_needed, episode_length_needed, env_reward_overview, experiment_name=None,
               out_dir=None):
    """Result file of the vary_hp evaluators (utils.py:144-160): same dict keys, so the reference's plot scripts load it."""
    import pandas as pd
    if experiment_name is None:
        experiment_name = "_experiment_"
    if out_dir is None:
        out_dir = os.getcwd()
    file_name = os.path.join(out_dir, str(mode) + '_' + experiment_name + '.pt')
    save_dict = {'config': config, 'reward_list': reward_list, 'train_steps_needed': train_steps_needed,
                 'episode_length_needed': episode_length_needed,
                 'env_reward_overview': pd.DataFrame.from_dict(env_reward_overview, orient="index")}
    torch.save(save_dict, file_name)
    return file_name
