"""GPU tests of the drop-in Python API (the reference's class / method surface) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def le():
    import learning_environments_b200 as pkg
    from learning_environments_b200 import agents, default_configs, envs, gtn, ops
    return dict(agents=agents, cfgs=default_configs, envs=envs, gtn=gtn, ops=ops)


def _small(le, name="cartpole_syn_env", **over):
    cfg = le["cfgs"].get(name)
    cfg["agents"]["ddqn"].update(dict(train_episodes=3, test_episodes=2, init_episodes=1, print_rate=10 ** 9), **over)
    return cfg


def test_virtual_env_step_matches_oracle_and_keeps_state(le):
    cfg = _small(le)
    torch.manual_seed(0)
    fac = le["envs"].EnvFactory(cfg)
    venv = fac.generate_virtual_env()
    theta = venv.env.theta().numpy()
    lane = le["agents"].le_config.lane_cfg(cfg, "ddqn", 0)
    s = venv.reset()
    assert s.shape == (4,) and s.dtype == torch.float32
    for a in (0, 1, 1, 0):
        s2, r, d = venv.step(torch.tensor([float(a)]))
        ons, orr, od = c_oracle.se_step(lane, theta, s.numpy(), a)
        assert rel_err(s2.numpy(), ons, 1e-2) < 1e-5 and rel_err(r.item(), orr, 1e-2) < 1e-5 and rel_err(d.item(), od, 1e-2) < 1e-5
        assert torch.equal(venv.env.state.cpu(), s2)           # self.state = next_state (envs/virtual_env.py:52)
        s = s2
    # batched entry with explicit states (envs/virtual_env.py:45-47)
    S = torch.rand(5, 4)
    A = torch.tensor([0.0, 1.0, 1.0, 0.0, 1.0])
    s2b, rb, db = venv.step(A, state=S)
    assert s2b.shape == (5, 4) and rb.shape == (5, 1) and db.shape == (5, 1)
    for i in range(5):
        ons, orr, od = c_oracle.se_step(lane, theta, S[i].numpy(), int(A[i]))
        assert rel_err(s2b[i].numpy(), ons, 1e-2) < 1e-5 and rel_err(rb[i].item(), orr, 1e-2) < 1e-5


def test_real_and_reward_env_wrappers(le):
    cfg = _small(le, "cartpole_reward_env")
    torch.manual_seed(1)
    fac = le["envs"].EnvFactory(cfg)
    real = fac.generate_real_env()
    real.env.seed(3)
    s = real.reset()
    assert s.shape == (4,) and float(s.abs().max()) <= 0.05
    total, steps, done = 0.0, 0, torch.tensor(0.0)
    while done < 0.5:
        s, r, done = real.step(real.get_random_action())
        total += float(r)
        steps += 1
    assert total == steps and 5 <= steps <= 200             # CartPole: reward 1 per step, random policy falls quickly
    renv = fac.generate_reward_env()
    renv.set_agent_params(same_action_num=1, gamma=0.99)
    renv.env.real_env.seed(5)
    s = renv.reset()
    lane = le["agents"].le_config.lane_cfg(cfg, "ddqn", 1, gamma=0.99)
    theta = renv.env.theta().numpy()
    s2, r, d = renv.step(torch.tensor([1.0]))
    want = c_oracle.rn_reward(lane, theta, s.numpy(), s2.numpy(), 1.0)
    assert rel_err(float(r), want, 1e-2) < 1e-5
    for t in (3, 101):
        c2 = _small(le, "cartpole_reward_env")
        c2["envs"]["CartPole-v0"]["reward_env_type"] = t
        c2["envs"]["CartPole-v0"]["info_dim"] = 2
        e = le["envs"].EnvFactory(c2).generate_reward_env()
        e.set_agent_params(1, 0.99)
        e.reset()
        with pytest.raises(ValueError, match="No info dict"):
            e.step(torch.tensor([0.0]))


def test_ddqn_train_and_test_match_oracle_lane(le):
    cfg = _small(le)
    torch.manual_seed(2)
    fac = le["envs"].EnvFactory(cfg)
    venv, real = fac.generate_virtual_env(), fac.generate_real_env()
    agent = le["agents"].select_agent(cfg, "DDQN")
    agent._seed = 99
    q0 = agent._theta.cpu().numpy().copy()
    rewards, lengths, rb = agent.train(env=venv, test_env=real)
    assert len(rewards) == len(lengths) == 3 and rb.get_size() == sum(lengths)
    lane = agent.last_run["cfg"]
    from learning_environments_b200.rng import lane_keys
    key = tuple(int(k) for k in lane_keys(99, 0, [0], [0], [0])[0])
    want = c_oracle.run_lane(lane, venv.env.theta().numpy(), key, q_init_w=q0)
    assert lengths == want["lengths"].tolist()
    assert np.allclose(rewards, want["rewards"])
    assert rel_err(agent._theta.cpu().numpy(), want["q_final"], 1e-2) < 1e-3
    assert not np.array_equal(agent.model.net[0].weight.detach().cpu().numpy().reshape(-1), q0[:228])   # module views the flat tensor
    # replay buffer returned by train(): first transition starts from the first reset state
    assert rb.state.shape[1] == 4 and float(rb.action[:rb.size].max()) <= 1.0
    # ... and it is THIS lane's ring (slot = lane id, warp-per-lane and multi-warp kernels alike), not uninitialised workspace:
    # inside the first episode every row continues the previous one, and every stored value is finite
    L0 = int(lengths[0])
    assert L0 >= 2 and np.array_equal(rb.next_state[:L0 - 1].cpu().numpy(), rb.state[1:L0].cpu().numpy())
    assert np.isfinite(rb.state[:rb.size].cpu().numpy()).all() and np.isfinite(rb.reward[:rb.size].cpu().numpy()).all()
    assert set(np.unique(rb.action[:rb.size].cpu().numpy())) <= {0.0, 1.0}
    test_rewards, test_lengths, _ = agent.test(env=real)
    assert len(test_rewards) == 2 and test_rewards == [float(l) for l in test_lengths]      # CartPole: reward == length
    assert set(agent.model.state_dict().keys()) == {"net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias"}


def test_ddqn_step_by_step_learn_matches_oracle(le):
    cfg = _small(le)
    torch.manual_seed(3)
    agent = le["agents"].select_agent(cfg, "DDQN")
    lane = agent._unit_cfg()
    th = agent._theta.cpu().numpy().copy()
    thT, m, v, t = th.copy(), np.zeros_like(th), np.zeros_like(th), 0
    rb = le["agents"].ReplayBuffer(state_dim=4, action_dim=1, device="cpu", max_size=1000)
    rng = np.random.RandomState(0)
    for i in range(300):
        rb.add(torch.from_numpy(rng.rand(4).astype(np.float32)), torch.tensor([float(rng.randint(2))]),
               torch.from_numpy(rng.rand(4).astype(np.float32)), torch.tensor(float(rng.rand())), torch.tensor(float(rng.rand() < 0.1)))
    real = le["envs"].EnvFactory(cfg).generate_real_env()
    for k in range(3):
        np.random.seed(10 + k)
        loss = agent.learn(replay_buffer=rb, env=real, episode=5)
        np.random.seed(10 + k)
        s, a, s2, r, d = rb.sample(agent.batch_size)
        rows = torch.cat([s, a, s2, r, d], dim=1).numpy()
        want, t = c_oracle.td_update(lane, th, thT, m, v, t, rows)
        assert rel_err(loss.item(), want) < 1e-5
    assert rel_err(agent._theta.cpu().numpy(), th, 1e-2) < 5e-5 and rel_err(agent._target.cpu().numpy(), thT, 1e-2) < 5e-5
    assert agent.it == 3
    st = torch.rand(4)
    q, a = c_oracle.q_forward(lane, th, st.numpy())
    assert int(agent.select_test_action(st, real)) == a
    agent.eps = 0.0
    assert int(agent.select_train_action(st, real, 0)) == a
    agent.update_parameters_per_episode(0)
    assert agent.eps == agent.eps_init
    agent.update_parameters_per_episode(1)
    assert agent.eps == max(agent.eps_init * agent.eps_decay, agent.eps_min)


def test_ddqn_vary_runs_with_sampled_hyperparameters(le):
    cfg = _small(le, train_episodes=2)
    le["agents"].DDQN_vary._rng = np.random.RandomState(4)
    fac = le["envs"].EnvFactory(cfg)
    venv, real = fac.generate_virtual_env(), fac.generate_real_env()
    kinds = set()
    for _ in range(6):
        agent = le["agents"].select_agent(cfg, "DDQN_vary")     # H in [19,171], hidden_layer in {1,2}: every sample must train
        lane, _ = agent._lane_cfg(venv, None, 2, False, 1e9)
        kinds.add(lane.q_is_register_resident())
        rewards, lengths, _ = agent.train(env=venv)            # test_env=None: virtual-env plateau rule
        assert len(rewards) == len(lengths) and len(rewards) >= 1
    assert kinds == {True, False}                               # both the register kernel and the general kernel were exercised


def test_gtn_master_on_gpu_matches_oracle_backed_master(le, tmp_path, monkeypatch):
    from tests.oracle_backend import OracleEvaluator, nes_update_numpy
    monkeypatch.chdir(tmp_path)
    cfg = _small(le, train_episodes=2)
    cfg["agents"]["gtn"].update(num_workers=4, max_iterations=2)
    torch.manual_seed(5)
    m = le["gtn"].GTN_Master(cfg, seed=31, verbose=False)
    theta0 = m.theta.clone()
    m.generation = 0
    m.evaluate_population()
    scores_gpu = (list(m.score_list), list(m.score_orig_list), list(m.sign_list))
    torch.manual_seed(5)
    ref = le["gtn"].GTN_Master(cfg, seed=31, evaluator_cls=OracleEvaluator, verbose=False)
    assert torch.equal(ref.theta, theta0)
    ref.generation = 0
    ref.evaluate_population()
    same = np.isclose(scores_gpu[0], ref.score_list) & np.isclose(scores_gpu[1], ref.score_orig_list)
    assert same.mean() >= 0.75                                  # chaos may move an occasional lane
    m.score_transform()
    m.update_env()
    assert not torch.equal(m.theta, theta0)
    if same.all() and scores_gpu[2] == ref.sign_list:
        ref.score_transform()
        th = ref.theta.clone()
        coef = torch.tensor([np.float32(ref.step_size * w) for w in ref.score_transform_list])
        # device noise (CUDA libm) vs numpy noise agree to 1 ulp of the normals -> theta to ~1e-7 relative
        nes_update_numpy(th, 4, 31, 0, ref.noise_std, ref.weight_decay, coef, torch.tensor(ref.sign_list))
        assert rel_err(m.theta.numpy(), th.numpy(), 1e-2) < 1e-5
    mean_score, mean_list, name = m.run()
    assert len(mean_list) == 2 and np.isfinite(mean_score)


def test_dueling_ddqn_agent_train_test(le):
    cfg = _small(le)
    cfg["agents"]["duelingddqn"].update(train_episodes=2, test_episodes=2, init_episodes=1, print_rate=10 ** 9)
    torch.manual_seed(7)
    fac = le["envs"].EnvFactory(cfg)
    venv, real = fac.generate_virtual_env(), fac.generate_real_env()
    agent = le["agents"].select_agent(cfg, "DuelingDDQN")
    assert set(agent.model.state_dict().keys()) == {"%s.%d.%s" % (s_, i, p) for s_ in ("feature_stream", "value_stream", "advantage_stream")
                                                    for i in (0, 2) for p in ("weight", "bias")}
    assert agent._theta.numel() == 11528
    agent._seed = 5
    q0 = agent._theta.cpu().numpy().copy()
    rewards, lengths, rb = agent.train(env=venv, test_env=real)
    lane = agent.last_run["cfg"]
    from learning_environments_b200.rng import lane_keys
    key = tuple(int(k) for k in lane_keys(5, 0, [0], [0], [0])[0])
    want = c_oracle.run_lane(lane, venv.env.theta().numpy(), key, q_init_w=q0)
    assert lengths[0] == want["lengths"][0] and len(rewards) == 2
    test_rewards, test_lengths, _ = agent.test(env=real)
    assert len(test_rewards) == 2 and test_rewards == [float(l) for l in test_lengths]
    # step-by-step learn() on the general unit kernel
    rbuf = le["agents"].ReplayBuffer(state_dim=4, action_dim=1, device="cpu", max_size=1000)
    rng = np.random.RandomState(0)
    for i in range(250):
        rbuf.add(torch.from_numpy(rng.rand(4).astype(np.float32)), torch.tensor([float(rng.randint(2))]),
                 torch.from_numpy(rng.rand(4).astype(np.float32)), torch.tensor(float(rng.rand())), torch.tensor(float(rng.rand() < 0.1)))
    th = agent._theta.cpu().numpy().copy()
    thT, m, v = agent._target.cpu().numpy().copy(), np.zeros_like(th), np.zeros_like(th)
    agent.reset_optimizer()
    np.random.seed(3)
    loss = agent.learn(replay_buffer=rbuf, env=real, episode=5)
    np.random.seed(3)
    s, a, s2, r, d = rbuf.sample(agent.batch_size)
    want_loss, _ = c_oracle.td_update(agent._unit_cfg(), th, thT, m, v, 0, torch.cat([s, a, s2, r, d], dim=1).numpy())
    assert rel_err(loss.item(), want_loss) < 1e-5


def test_vary_hp_batched_agents_match_oracle_per_lane(le):
    """BASELINE config 4 in miniature: agents with per-lane lr / batch_size / hidden_size / hidden_layer on one SE."""
    from learning_environments_b200 import vary_hp
    from learning_environments_b200.rng import lane_keys
    cfg = le["cfgs"].get("cartpole_syn_env")
    torch.manual_seed(11)
    venv = le["envs"].EnvFactory(cfg).generate_virtual_env()
    theta = venv.env.theta().numpy()
    over = dict(print_rate=10, early_out_num=2, train_episodes=3, init_episodes=1, test_episodes=2, early_out_virtual_diff=0.01)
    n = 12
    rewards, steps, episodes, cfgs = vary_hp.train_test_agents(cfg, theta, agents_num=n, seed=3, overrides=over)
    assert len({(c.q_hidden, c.batch_size, c.q_layers) for c in cfgs}) > 6          # genuinely heterogeneous lanes
    assert {c.q_is_register_resident() for c in cfgs} == {True, False}
    keys = lane_keys(3, 0, np.zeros(n, int), np.zeros(n, int), np.arange(n))
    ok = 0
    for i in range(n):
        want = c_oracle.run_lane(cfgs[i], theta, tuple(int(k) for k in keys[i]))
        assert len(rewards[i]) == 2 and episodes[i] >= 1
        if steps[i] == want["train_steps"]:
            ok += 1
            assert episodes[i] == want["n_episodes"]
    assert ok >= 0.7 * n


def test_run_vary_hp_all_models_in_one_launch(le, tmp_path):
    """experiments/syn_env_run_vary_hp.py modes 0 and 2 on GTN-format checkpoints: result file keys / shapes of utils.save_lists."""
    from learning_environments_b200 import vary_hp
    cfg = le["cfgs"].get("cartpole_syn_env")
    cfg["agents"]["ddqn_vary"]["vary_hp"] = True
    model_dir = tmp_path / "GTN_models_CartPole-v0"
    model_dir.mkdir()
    thetas = {}
    for i, suffix in enumerate(["ZZZAAA", "AAAZZZ", "MMMMMM"]):
        torch.manual_seed(100 + i)
        venv = le["envs"].EnvFactory(cfg).generate_virtual_env()
        name = "CartPole-v0_%s.pt" % suffix
        torch.save({"model": venv.state_dict(), "config": cfg}, str(model_dir / name))
        thetas[name] = venv.env.theta().numpy()
    over = dict(print_rate=10, early_out_num=2, train_episodes=4, init_episodes=1, test_episodes=3, early_out_virtual_diff=0.01)
    f2 = vary_hp.run_vary_hp(mode=2, experiment_name="t", model_num=2, agents_num=3, model_dir=str(model_dir), env_name="CartPole",
                             out_dir=str(tmp_path), seed=5, overrides=over)
    d = torch.load(f2, weights_only=False)
    assert os.path.basename(f2) == "2_t.pt"
    assert len(d["reward_list"]) == 6 and all(len(r) == 3 for r in d["reward_list"])
    assert len(d["train_steps_needed"]) == 6 and len(d["episode_length_needed"]) == 6
    assert list(d["env_reward_overview"].index) == ["CartPole-v0_AAAZZZ.pt", "CartPole-v0_MMMMMM.pt"]     # sorted by suffix
    assert d["env_reward_overview"].shape == (2, 9)
    # per-lane check of the first model's agents against the CPU restatement
    from learning_environments_b200.rng import lane_keys
    rng = np.random.RandomState(5)
    cfgs = vary_hp.sample_agent_cfgs(cfg, 6, rng, over, True, None)
    keys = lane_keys(5, 0, np.arange(6) // 3, np.zeros(6, int), np.arange(6) % 3)
    ok = 0
    for i in range(6):
        want = c_oracle.run_lane(cfgs[i], thetas[d["env_reward_overview"].index[i // 3]], tuple(int(k) for k in keys[i]))
        ok += int(d["train_steps_needed"][i][0] == want["train_steps"] and d["episode_length_needed"][i][0] == want["n_episodes"])
    assert ok >= 4
    f0 = vary_hp.run_vary_hp(mode=0, experiment_name="t", model_num=2, agents_num=2, model_dir=str(model_dir), env_name="CartPole",
                             out_dir=str(tmp_path), seed=6, overrides=over)
    d0 = torch.load(f0, weights_only=False)
    assert len(d0["reward_list"]) == 4 and list(d0["env_reward_overview"].index) == ["CartPole-v0_0", "CartPole-v0_1"]
    assert all(1 <= e[0] <= 4 for e in d0["episode_length_needed"])


def test_integration_md_ctypes_stub_runs_as_written(le):
    """The reference-side binding shown in INTEGRATION.md (section 2) is executed verbatim: host buffers in, scores out."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(.*?)```", text, re.S).group(1)
    ns = {}
    cwd = os.getcwd()
    os.chdir(root)                      # the stub loads the library by its repo-relative path
    try:
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
        cfg = _small(le)
        lane = le["agents"].le_config.lane_cfg(cfg, "ddqn", 0)
        stub_cfg = ns["LaneCfg"].from_buffer_copy(bytes(lane))
        assert ns["lib"].le_sizeof_lane_cfg() == len(bytes(lane))
        torch.manual_seed(2)
        theta = le["envs"].EnvFactory(cfg).generate_virtual_env().env.theta().numpy()[None].copy()
        keys = np.array([[5, 6], [7, 8], [9, 10]], np.uint32)
        scores, rewards, lengths = ns["calc_scores"](stub_cfg, theta, np.zeros(3, np.int32), keys)
    finally:
        os.chdir(cwd)
    for i in range(3):
        want = c_oracle.run_lane(lane, theta[0], (int(keys[i, 0]), int(keys[i, 1])))
        assert lengths[i, 0] == want["lengths"][0]                     # pre-learning episode: exact
    assert np.isfinite(scores).all() and (lengths[:, 0] > 0).all()
