"""TEST INFRASTRUCTURE ONLY — numpy restatement of the Philox4x32-10 streams the CUDA kernels use.

The reference seeds python/numpy/torch generators from the wall clock (agents/GTN_worker.py:24-25,
experiments/GTNC_evaluate_cartpole.py:83-95), so there is no reference stream to match; the build replaces
every random source of the hot path by counter-based Philox4x32-10 (Salmon et al., SC'11; Random123
constants) so that the CPU oracle, the reference under RNG injection and the GPU kernels consume identical
words.  Known-answer vectors of Random123's kat_vectors pin the block function (tests/test_philox.py).

Stream layout — per *lane* (one agent = one (member, variant, eval) of the NES population) a 64-bit key
(k0, k1); counter = (c0, c1, purpose, 0):

  P_ACT         (train_step, 0)        w0 -> u = (w0>>8)*2^-24 ; explore iff u < eps (double compare)
                                       w1 -> random action = mulhi(w1, action_dim)
  P_SAMPLE      (learn_iter, j)        replay indices 4j..4j+3 = mulhi(w_k, size)      (utils.py:35)
  P_RESET_TRAIN (episode, 0)           4 uniforms for the training env's reset draw
  P_RESET_TEST  (test_call, test_ep)   4 uniforms for a test-episode reset draw
  P_QINIT       (block, 0)             Q-net init U(-1/sqrt(fan_in), +1/sqrt(fan_in)), params 4*block..
  P_NOISE       (block, member)        NES perturbation normals (key = (seed, generation)), Box-Muller in fp64

  P_TD3_EXPO    (c0, phase) + block    Exp(1) draws of F.gumbel_softmax (TD3_discrete groundwork): e = -ln((w+0.5)*2^-32) in fp64 -> f32;
                                       phase 0: train action (c0 = train_step), 1: test action (c0 = test_call << 16 | step in the episode, episode in
                                       bits 16.. of the 4th counter word: test episodes are independent, so they can run in parallel),
                                       2: target policy in learn (c0 = learn_iter), 3: policy update in learn; the 4th counter word
                                       is the block index (draws 4*block .. 4*block+3, row-major over [B][action_dim])
  P_TD3_NORMAL  (c0, phase) + block    N(0,1) draws (Box-Muller as P_NOISE): phase 0 train action noise, 1 test action noise,
                                       2 policy noise of learn (randn_like(actions))

uniform in [lo,hi):  lo + (hi-lo) * (w * 2^-32)   evaluated in float64.
"""
import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

P_ACT = 1
P_SAMPLE = 2
P_RESET_TRAIN = 3
P_RESET_TEST = 4
P_QINIT = 5
P_NOISE = 6
P_HP = 7
P_TD3_EXPO = 8
P_TD3_NORMAL = 9


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Scalar Philox4x32-R on python ints. Returns a tuple of four uint32."""
    c0 &= MASK; c1 &= MASK; c2 &= MASK; c3 &= MASK; k0 &= MASK; k1 &= MASK
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return (c0, c1, c2, c3)


def philox4x32_np(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R: counters are broadcastable uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & np.uint64(MASK) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(int(k0) & MASK)
    k1 = np.uint64(int(k1) & MASK)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(rounds):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> s32, p0 & m
        hi1, lo1 = p1 >> s32, p1 & m
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m, lo1, (hi0 ^ c3 ^ k1) & m, lo0
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def mulhi(w, n):
    return (int(w) * int(n)) >> 32


def sample_indices(key, learn_iter, batch_size, size):
    """Replay sample of utils.py:35 (`np.random.randint(0, size, B)`), restated on the P_SAMPLE stream."""
    nblk = (batch_size + 3) // 4
    w = philox4x32_np(np.uint64(learn_iter), np.arange(nblk, dtype=np.uint64), P_SAMPLE, 0, key[0], key[1])
    w = w.reshape(-1)[:batch_size].astype(np.uint64)
    return ((w * np.uint64(size)) >> np.uint64(32)).astype(np.int32)


def uniform_f64(words, lo, hi):
    w = np.asarray(words, dtype=np.float64)
    return lo + (hi - lo) * (w * (1.0 / 4294967296.0))


def qnet_init(key, n_params, bounds):
    """bounds: float64 array [n_params] of 1/sqrt(fan_in) per parameter. Returns float32 [n_params]."""
    nblk = (n_params + 3) // 4
    w = philox4x32_np(np.arange(nblk, dtype=np.uint64), 0, P_QINIT, 0, key[0], key[1]).reshape(-1)[:n_params]
    u = (w.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    return ((2.0 * u - 1.0) * np.asarray(bounds, dtype=np.float64)).astype(np.float32)


def normals(seed, generation, member, n):
    """n standard normals (float32) of the NES perturbation stream for `member` in `generation`.

    Box-Muller in float64 on word pairs: u1 = (w0+1)*2^-32 in (0,1], u2 = w1*2^-32;
    z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = sqrt(-2 ln u1) sin(2 pi u2); block b gives normals 4b..4b+3
    = (z0(w0,w1), z1(w0,w1), z0(w2,w3), z1(w2,w3)); results rounded to float32.
    """
    nblk = (n + 3) // 4
    w = philox4x32_np(np.arange(nblk, dtype=np.uint64), np.uint64(member), P_NOISE, 0, seed, generation)
    w = w.astype(np.float64)
    out = np.empty((nblk, 4), dtype=np.float64)
    for a, b, o in ((0, 1, 0), (2, 3, 2)):
        u1 = (w[:, a] + 1.0) * (1.0 / 4294967296.0)
        u2 = w[:, b] * (1.0 / 4294967296.0)
        r = np.sqrt(-2.0 * np.log(u1))
        t = (2.0 * np.pi) * u2
        out[:, o] = r * np.cos(t)
        out[:, o + 1] = r * np.sin(t)
    return out.reshape(-1)[:n].astype(np.float32)


def lane_key(seed, generation, member, variant, eval_idx=0):
    """Derives the 64-bit lane key from (seed, generation, member, variant, eval): one Philox call keyed by
    (seed, 0x4C414E45 'LANE') on counter (generation, member, variant, eval_idx)."""
    w = philox4x32(generation, member, variant, eval_idx, seed, 0x4C414E45)
    return (w[0], w[1])


def td3_expo(key, phase, c0, n, sub=0):
    """n Exp(1) draws (float32) of the P_TD3_EXPO stream (sub: bits 16.. of the 4th counter word, the test episode)."""
    nblk = (n + 3) // 4
    w = philox4x32_np(np.uint64(c0), np.uint64(phase), P_TD3_EXPO, np.arange(nblk, dtype=np.uint64) + np.uint64(sub << 16), key[0], key[1]).reshape(-1)[:n]
    u = (w.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    return (-np.log(u)).astype(np.float32)


def td3_normal(key, phase, c0, n, sub=0):
    """n N(0,1) draws (float32) of the P_TD3_NORMAL stream (same Box-Muller as normals())."""
    nblk = (n + 3) // 4
    w = philox4x32_np(np.uint64(c0), np.uint64(phase), P_TD3_NORMAL, np.arange(nblk, dtype=np.uint64) + np.uint64(sub << 16), key[0], key[1]).astype(np.float64)
    out = np.empty((nblk, 4), dtype=np.float64)
    for a, b, o in ((0, 1, 0), (2, 3, 2)):
        u1 = (w[:, a] + 1.0) * (1.0 / 4294967296.0)
        u2 = w[:, b] * (1.0 / 4294967296.0)
        r = np.sqrt(-2.0 * np.log(u1))
        t = (2.0 * np.pi) * u2
        out[:, o] = r * np.cos(t)
        out[:, o + 1] = r * np.sin(t)
    return out.reshape(-1)[:n].astype(np.float32)
