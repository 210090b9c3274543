"""Helpers shared by the TD3-update parity tests (CPU oracle and GPU)."""
import numpy as np

NETS = ("actor", "critic0", "critic1", "actor_target", "critic0_target", "critic1_target")


def nets_from(g, prefix):
    return {name: [np.asarray(g[f"{prefix}_{name}_{i}"]) for i in range(6)] for name in NETS}


def make_oracle(T, g):
    n = nets_from(g, "init")
    gamma, tau, delay, _sigma, clip, lr = [float(x) for x in g["hyper"]]
    return T.TD3UpdateOracle(n["actor"], [n["critic0"], n["critic1"]], n["actor_target"], [n["critic0_target"], n["critic1_target"]], lr=lr,
                             gamma=gamma, tau=tau, policy_delay=int(delay), target_noise_clip=clip)


def replay(o, g):
    for k in range(g["noise"].shape[0]):
        o.step(g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k], g["noise"][k])
    return {"actor": o.actor, "critic0": o.critics[0], "critic1": o.critics[1], "actor_target": o.actor_target,
            "critic0_target": o.critic_targets[0], "critic1_target": o.critic_targets[1]}


def random_nets(rng, h1, h2):
    """torch nn.Linear default init (kaiming-uniform a=sqrt(5) -> U(+-1/sqrt(fan_in)) for weights and biases)."""
    def mlp(i, o):
        out = []
        for fi, fo in ((i, h1), (h1, h2), (h2, o)):
            b = 1.0 / np.sqrt(fi)
            out += [rng.uniform(-b, b, (fo, fi)).astype(np.float32), rng.uniform(-b, b, fo).astype(np.float32)]
        return out
    a, c0, c1 = mlp(4, 2), mlp(6, 1), mlp(6, 1)
    cp = lambda ps: [t.copy() for t in ps]  # noqa: E731
    return {"actor": a, "critic0": c0, "critic1": c1, "actor_target": cp(a), "critic0_target": cp(c0), "critic1_target": cp(c1)}


SAC_NETS = ("actor", "critic0", "critic1", "critic0_target", "critic1_target")


def sac_nets_from(g, prefix):
    return {name: [np.asarray(g[f"{prefix}_{name}_{i}"]) for i in range(6)] for name in SAC_NETS}


def make_sac_oracle(T, g):
    n = sac_nets_from(g, "init")
    gamma, tau, target_entropy, lr, interval = [float(x) for x in g["hyper"]]
    return T.SACUpdateOracle(n["actor"], [n["critic0"], n["critic1"]], [n["critic0_target"], n["critic1_target"]], lr=lr, gamma=gamma, tau=tau,
                             target_entropy=target_entropy, log_ent_coef=float(g["init_log_ent_coef"][0]), target_update_interval=int(interval))


def replay_sac(o, g):
    for k in range(g["eps_pi"].shape[0]):
        o.step(g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k], g["eps_pi"][k], g["eps_next"][k])
    return {"actor": o.actor, "critic0": o.critics[0], "critic1": o.critics[1], "critic0_target": o.critic_targets[0], "critic1_target": o.critic_targets[1]}


def random_sac_nets(rng, h1, h2):
    n = random_nets(rng, h1, h2)
    b = 1.0 / np.sqrt(h2)
    n["actor"][4] = rng.uniform(-b, b, (4, h2)).astype(np.float32)  # [mu; log_std] head
    n["actor"][5] = rng.uniform(-b, b, 4).astype(np.float32)
    return {k: n[k] for k in SAC_NETS}


BCQ_NETS = ("vae_enc", "vae_dec", "pert", "critic0", "critic1", "vae_enc_target", "vae_dec_target", "pert_target", "critic0_target", "critic1_target")


def bcq_nets_from(g, prefix):
    return {name: [np.asarray(g[f"{prefix}_{name}_{i}"]) for i in range(6)] for name in BCQ_NETS}


def make_bcq_oracle(T, g):
    n = bcq_nets_from(g, "init")
    gamma, tau, phi, lr, delay = [float(x) for x in g["hyper"]]
    return T.BCQUpdateOracle(n["vae_enc"], n["vae_dec"], n["pert"], [n["critic0"], n["critic1"]], lr=lr, gamma=gamma, tau=tau,
                             max_perturbation=phi, actor_delay=int(delay))


def replay_bcq(o, g):
    for k in range(g["eps_vae"].shape[0]):
        o.step(g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k], g["eps_vae"][k],
               g["z_next"][k], g["z_actor"][k])
    return {"vae_enc": o.vae_enc, "vae_dec": o.vae_dec, "pert": o.pert, "critic0": o.critics[0], "critic1": o.critics[1],
            "vae_enc_target": o.vae_enc, "vae_dec_target": o.vae_dec,  # the target VAE is a copy of the VAE (bcq.py:158-159)
            "pert_target": o.pert_target, "critic0_target": o.critic_targets[0], "critic1_target": o.critic_targets[1]}


MA_NETS = tuple(f"{n}{t}" for n in ("actor0", "actor1", "critic0_0", "critic0_1", "critic1_0", "critic1_1") for t in ("", "_target"))


def ma_nets_from(g, prefix):
    return {name: [np.asarray(g[f"{prefix}_{name}_{i}"]) for i in range(6)] for name in MA_NETS}


def make_ma_oracle(T, g, centralised):
    n = ma_nets_from(g, "init")
    gamma, tau, delay, clip, lr0, lr1 = [float(x) for x in g["hyper"]]
    return T.MultiAgentDDPGOracle([n["actor0"], n["actor1"]], [[n["critic0_0"], n["critic0_1"]], [n["critic1_0"], n["critic1_1"]]],
                                  [[0, 1], [2, 3]], [[0], [1]], centralised, *T.MultiAgentDDPGOracle.reference_lrs([lr0, lr1]), gamma=gamma, tau=tau, policy_delay=int(delay),
                                  target_noise_clip=clip)


def replay_ma(o, g):
    for k in range(g["noise"].shape[0]):
        o.step(g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k], list(g["noise"][k]))
    out = {}
    for i in range(2):
        out[f"actor{i}"], out[f"actor{i}_target"] = o.actors[i], o.actor_targets[i]
        for k in range(2):
            out[f"critic{i}_{k}"], out[f"critic{i}_{k}_target"] = o.critics[i][k], o.critic_targets[i][k]
    return out
