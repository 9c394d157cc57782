"""Device-resident replay buffer: the reference's `ReplayBuffer` / `EvaluationDataset`
(ikea/data_utils/replay_buffer.py:26-127) + the `DataLoader(shuffle=True)` that feeds `train_step`
(ikea/training/trainSQN.py), re-designed for the native path.

The reference keeps the columns in numpy, indexes one row per `__getitem__`, collates 256 of them in python per step
and ships the batch to the GPU: ~10^4-10^5 sessions/s.  Here the columns are uploaded ONCE, an epoch is one
permutation (the same `torch.randperm(n, generator=g)` the DataLoader's RandomSampler draws, so a seeded epoch visits
the rows in the reference's order), and every batch is one launch of `rec_gather_batch` (include/recsys_b200.h)
that writes the tuple `(s, a, r, s_next, true_len, true_next_len, is_end)` straight into HBM -- no host work and no
host->device copy per step.  Constructor keywords, `__len__` and `__getitem__` are the reference's.
"""

import ctypes as C

import numpy as np
import torch

from ... import _native as N


def _read_json_lines(path, names):
    import pandas as pd
    df = pd.read_json(path, orient="records", lines=True)
    return {k: df[v] for k, v in names.items()}


class DeviceReplayBuffer:
    """Drop-in for `ReplayBuffer` (ikea/data_utils/replay_buffer.py:26-75) with device-side sampling."""

    COLUMNS = ("states", "actions", "reward", "next_states", "true_state_len", "true_next_state_len", "is_end")

    def __init__(self, dir=None, state_name="state", next_state_name="next_state", action_name="action",
                 end_name="is_end", state_len_name="true_state_len", true_next_state_len_name="true_next_state_len",
                 reward_name="r_act", arrays=None):
        self.dir = dir
        self.state_name, self.next_state_name = state_name, next_state_name
        self.reward_name, self.action_name, self.end_name = reward_name, action_name, end_name
        if arrays is None:
            cols = _read_json_lines(dir, dict(states=state_name, next_states=next_state_name, actions=action_name,
                                              reward=reward_name, true_state_len=state_len_name,
                                              true_next_state_len=true_next_state_len_name, is_end=end_name))
            arrays = dict(states=np.array(cols["states"].values.tolist()),
                          next_states=np.array(cols["next_states"].values.tolist()),
                          actions=cols["actions"].to_numpy(), reward=cols["reward"].to_numpy(),
                          true_state_len=cols["true_state_len"].to_numpy(),
                          true_next_state_len=cols["true_next_state_len"].to_numpy(), is_end=cols["is_end"].to_numpy())
        missing = [c for c in self.COLUMNS if c not in arrays]
        if missing:
            raise ValueError(f"replay buffer columns missing: {missing}")
        for c in self.COLUMNS:  # numpy, like the reference (host-side indexing keeps working)
            setattr(self, c, np.asarray(arrays[c]))
        n = len(self.actions)
        if self.states.ndim != 2 or self.next_states.shape != self.states.shape or any(
                len(getattr(self, c)) != n for c in self.COLUMNS):
            raise ValueError("replay buffer columns disagree in length / shape")
        self._dev = None
        self._engine = None
        self._n, self._L = n, int(self.states.shape[1])

    @classmethod
    def from_arrays(cls, states, actions, reward, next_states, true_state_len, true_next_state_len, is_end):
        return cls(arrays=dict(states=states, actions=actions, reward=reward, next_states=next_states,
                               true_state_len=true_state_len, true_next_state_len=true_next_state_len, is_end=is_end))

    @classmethod
    def from_event_log(cls, engine, session_ids, item_ids, pad_id, pad_pos="end", rewards=None):
        """Build the buffer ON THE DEVICE from a raw event log sorted by session: what
        `preprocess_train_data_incl_act_rew` (recommenders/data_utils/preprocessing.py:199-268) computes with a pandas
        groupby-apply per session, as one launch of `rec_build_replay_rows`.  `engine` supplies state_size, the device
        and the stream.  The host only derives the CSR session offsets (one vectorised numpy pass)."""
        sid = np.asarray(session_ids)
        n = len(sid)
        if n == 0:
            raise ValueError("empty event log")
        starts = np.flatnonzero(np.concatenate(([True], sid[1:] != sid[:-1])))
        off = np.concatenate((starts, [n])).astype(np.int64)
        dev, L = engine.device, int(engine.cfg["state_size"])
        d_off = torch.from_numpy(off).to(dev)
        d_items = torch.as_tensor(np.ascontiguousarray(item_ids)).to(torch.int64).to(dev)
        d_rew = None if rewards is None else torch.as_tensor(np.ascontiguousarray(rewards)).to(torch.float32).to(dev)
        i64 = dict(dtype=torch.int64, device=dev)
        cols = dict(s=torch.empty(n, L, **i64), a=torch.empty(n, **i64), r=torch.zeros(n, dtype=torch.float32, device=dev),
                    sn=torch.empty(n, L, **i64), ln=torch.empty(n, **i64), nl=torch.empty(n, **i64),
                    e=torch.empty(n, dtype=torch.uint8, device=dev))
        out = engine._batch(n, cols["s"], cols["a"], cols["ln"], cols["r"], cols["sn"], cols["nl"], cols["e"])
        N.check(engine.lib, engine.handle,
                engine.lib.rec_build_replay_rows(engine.handle, C.c_void_p(d_off.data_ptr()), len(off) - 1,
                                                 C.c_void_p(d_items.data_ptr()),
                                                 C.c_void_p(d_rew.data_ptr() if d_rew is not None else 0), n, int(pad_id),
                                                 1 if pad_pos == "end" else 0, C.byref(out)), "rec_build_replay_rows")
        torch.cuda.synchronize(dev)
        self = cls.__new__(cls)
        self.dir = None
        self._dev, self._engine = cols, engine
        self._n, self._L = n, L
        return self

    def _host(self, name):
        """Host mirror of a column of a device-built buffer (lazy; the training path never needs it)."""
        key = dict(states="s", actions="a", reward="r", next_states="sn", true_state_len="ln",
                   true_next_state_len="nl", is_end="e")[name]
        v = self._dev[key].cpu().numpy()
        return v.astype(bool) if name == "is_end" else v

    def __getattr__(self, name):
        if name in DeviceReplayBuffer.COLUMNS and "_dev" in self.__dict__ and self.__dict__["_dev"] is not None:
            v = self._host(name)
            self.__dict__[name] = v
            return v
        raise AttributeError(name)

    # ---- the reference's Dataset protocol ---------------------------------------------------------------
    def __len__(self):
        return self._n

    def __getitem__(self, idx):
        return (self.states[idx], self.actions[idx], self.reward[idx], self.next_states[idx], self.true_state_len[idx],
                self.true_next_state_len[idx], self.is_end[idx])

    # ---- device residency ----------------------------------------------------------------------------------
    def to_device(self, device):
        """Upload every column once: int64 ids / lengths, float32 rewards, uint8 is_end."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DeviceReplayBuffer.to_device needs a CUDA device (there is no CPU path)")
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dtype=dt).to(device).contiguous()
        self._dev = dict(s=t(self.states, torch.int64), a=t(self.actions, torch.int64), r=t(self.reward, torch.float32),
                         sn=t(self.next_states, torch.int64), ln=t(self.true_state_len, torch.int64),
                         nl=t(self.true_next_state_len, torch.int64), e=t(self.is_end.astype(np.uint8), torch.uint8))
        return self

    @property
    def device(self):
        return None if self._dev is None else self._dev["s"].device

    def bytes_on_device(self):
        return 0 if self._dev is None else sum(v.numel() * v.element_size() for v in self._dev.values())

    @staticmethod
    def epoch_permutation(n, shuffle=True, generator=None):
        """Row order of one epoch: exactly what `iter(DataLoader(buffer, shuffle=shuffle, generator=generator))` visits.

        A DataLoader iterator first draws its `_base_seed` from the generator (or the global RNG), then the
        RandomSampler draws ONE `torch.randperm(n)` -- from `generator`, or, without one, from a fresh generator seeded
        by another draw of the global RNG (torch/utils/data/dataloader.py, sampler.py).  Both draws are replayed here,
        so `torch.manual_seed(seed)` + this method visits the rows in the reference loop's order."""
        if not shuffle:
            return torch.arange(n)
        torch.empty((), dtype=torch.int64).random_(generator=generator)  # the iterator's _base_seed
        if generator is None:
            generator = torch.Generator()
            generator.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
        return torch.randperm(n, generator=generator)

    @staticmethod
    def batch_bounds(n, batch_size, drop_last=False, rank=0, world=1):
        """[(lo, hi)) slices of the epoch order, like `BatchSampler`.

        world > 1 (vocabulary-sharded multi-GPU runs: every rank contributes `batch_size` local sessions to a global
        batch of `world * batch_size`): rank r gets the r-th `batch_size` slice of every GLOBAL batch, so that the
        concatenation over ranks is exactly the single-process batch of size `world * batch_size` -- the G-GPU run
        trains on the same global batches, in the same order, as a 1-GPU run with the global batch size.  Ragged
        global batches are dropped (every rank must take the same number of steps with the same local size)."""
        if world == 1:
            stop = n - n % batch_size if drop_last else n
            return [(lo, min(lo + batch_size, stop)) for lo in range(0, stop, batch_size)]
        if not 0 <= rank < world:
            raise ValueError(f"rank {rank} outside world {world}")
        g = batch_size * world
        return [(lo + rank * batch_size, lo + (rank + 1) * batch_size) for lo in range(0, n - n % g, g)]

    def batches(self, engine, batch_size, shuffle=True, generator=None, drop_last=False, slots=4, rank=0, world=1):
        """Yield `(s, a, r, s_next, true_len, true_next_len, is_end)` device tensors for one epoch.

        `engine`: the trainer's native engine (`trainer._ready(batch_size)`); the gather runs on its stream, so a
        `train_step_async(*batch)` issued next is ordered after it.  `slots` output buffers rotate: a yielded batch
        stays valid until `slots - 1` later batches have been requested."""
        if self._dev is None:
            raise RuntimeError("call to_device(device) first")
        d = self._dev
        n, L, dev = self._n, self._L, d["s"].device
        perm = self.epoch_permutation(n, shuffle, generator).to(dev)
        cols = engine._batch(n, d["s"], d["a"], d["ln"], d["r"], d["sn"], d["nl"], d["e"])
        ring = []
        for _ in range(max(2, slots)):
            i64 = dict(dtype=torch.int64, device=dev)
            ring.append((torch.empty(batch_size, L, **i64), torch.empty(batch_size, **i64),
                         torch.empty(batch_size, dtype=torch.float32, device=dev), torch.empty(batch_size, L, **i64),
                         torch.empty(batch_size, **i64), torch.empty(batch_size, **i64),
                         torch.empty(batch_size, dtype=torch.uint8, device=dev)))
        # (world > 1: every rank must pass a generator with the same seed -- the permutation is replicated, not sent)
        for k, (lo, hi) in enumerate(self.batch_bounds(n, batch_size, drop_last, rank, world)):
            B = hi - lo
            s, a, r, sn, ln, nl, e = (x[:B] for x in ring[k % len(ring)])
            out = engine._batch(B, s, a, ln, r, sn, nl, e)
            N.check(engine.lib, engine.handle,
                    engine.lib.rec_gather_batch(engine.handle, C.byref(cols), n, C.c_void_p(perm[lo:hi].data_ptr()), B,
                                                C.byref(out)), "rec_gather_batch")
            yield s, a, r, sn, ln, nl, e


class DeviceEvaluationDataset:
    """Drop-in for `EvaluationDataset` (ikea/data_utils/replay_buffer.py:78-127): `(s, a, s_len)` only."""

    def __init__(self, dir=None, state_name="state", action_name="action", state_len_name="true_state_len", arrays=None):
        self.dir, self.state_name, self.action_name = dir, state_name, action_name
        if arrays is None:
            cols = _read_json_lines(dir, dict(states=state_name, actions=action_name, true_state_len=state_len_name))
            arrays = dict(states=np.array(cols["states"].values.tolist()), actions=cols["actions"].to_numpy(),
                          true_state_len=cols["true_state_len"].to_numpy())
        self.states = np.asarray(arrays["states"])
        self.actions = np.asarray(arrays["actions"])
        self.true_state_len = np.asarray(arrays["true_state_len"])
        self._dev = None

    def __len__(self):
        return len(self.actions)

    def __getitem__(self, idx):
        return (self.states[idx], self.actions[idx], self.true_state_len[idx])

    def to_device(self, device):
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dtype=torch.int64).to(device).contiguous()
        self._dev = (t(self.states), t(self.actions), t(self.true_state_len))
        return self

    def batches(self, batch_size):
        """Sequential `(s, a, s_len)` device batches (views of the resident columns: evaluation never shuffles)."""
        if self._dev is None:
            raise RuntimeError("call to_device(device) first")
        s, a, ln = self._dev
        for lo in range(0, len(self), batch_size):
            yield s[lo:lo + batch_size], a[lo:lo + batch_size], ln[lo:lo + batch_size]
