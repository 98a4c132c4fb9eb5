"""Configuration namespaces with the reference's key names.

The reference builds an argparse.Namespace by dict-unioning YAML sections (launch.py:24-41) and
passes it to Dreamer(conf) (dreamer.py:21).  The drop-in keeps that contract: any object with these
attributes works, including one produced by the reference's own launcher.  The values below restate
the hot-path keys of config/defaults.yaml (`defaults`:1-120, `atari`:188-201, `dmc`:203-213) so that
benchmarks and tests do not need the reference checkout at run time.
"""
from argparse import Namespace

# config/defaults.yaml `defaults` section, hot-path keys only
DEFAULTS = dict(
    # features
    image_key="image", image_size=64, image_channels=3, image_categorical=False, action_dim=0, clip_rewards=None,
    map_key=None, map_size=0, map_channels=0, map_categorical=True, goals_size=0,
    # training
    reset_interval=200, iwae_samples=1, kl_balance=0.8, kl_weight=1.0, image_weight=1.0, vecobs_weight=1.0,
    reward_weight=1.0, terminal_weight=1.0, adam_lr=3.0e-4, adam_lr_actor=1.0e-4, adam_lr_critic=1.0e-4,
    adam_eps=1.0e-5, keep_state=True, batch_length=48, batch_size=32, device="cuda:0", grad_clip=200,
    grad_clip_ac=200, image_decoder_min_prob=0, amp=False, probe_gradients=False,
    # model
    model="dreamer", deter_dim=2048, stoch_dim=32, stoch_discrete=32, hidden_dim=1000, gru_layers=1, gru_type="gru",
    layer_norm=True, vecobs_size=0, image_encoder="cnn", cnn_depth=48, image_encoder_layers=0, image_decoder="cnn",
    image_decoder_layers=0, reward_input=False, reward_decoder_layers=4, reward_decoder_categorical=None,
    terminal_decoder_layers=4,
    # probe
    probe_model="none", map_decoder="dense", map_hidden_layers=4, map_hidden_dim=1024,
    # actor critic
    gamma=0.995, lambda_gae=0.95, entropy=0.003, target_interval=100, imag_horizon=15, actor_grad="reinforce",
    actor_dist="onehot",
    # auxiliary critic
    aux_critic=False, aux_critic_weight=1.0, gamma_aux=0.99, lambda_gae_aux=0.95, target_interval_aux=1000,
)

SECTIONS = {
    # config/defaults.yaml:188-201
    "atari": dict(action_dim=18, clip_rewards="tanh", deter_dim=1024, kl_weight=0.1, gamma=0.99, entropy=0.001),
    # config/defaults.yaml:203-213
    "dmc": dict(action_dim=12, entropy=1.0e-4, actor_grad="dynamics", actor_dist="tanh_normal", clip_rewards="tanh"),
}

# BASELINE.json configs (the shipped YAML differs from BASELINE's wording, so the
# overrides are explicit).  DMC: the shipped actor_grad=dynamics asserts at a2c.py:131 upstream
#  -> graded variant is reinforce + tanh_normal.
PRESETS = {
    "atari": (("atari",), dict(deter_dim=2048, batch_size=50, batch_length=50)),
    "atari_iwae": (("atari",), dict(deter_dim=2048, batch_size=50, batch_length=50, iwae_samples=4)),
    "dmc": (("dmc",), dict(deter_dim=1024, batch_size=50, batch_length=50, actor_grad="reinforce")),
    "atari_shipped": (("atari",), dict()),
    # small shapes for tests / smoke (same structure, every dimension shrunk)
    "tiny": (("atari",), dict(deter_dim=64, stoch_dim=4, stoch_discrete=8, hidden_dim=40, cnn_depth=4, action_dim=5,
                              batch_size=3, batch_length=4, imag_horizon=3)),
    "tiny_dmc": (("dmc",), dict(deter_dim=64, stoch_dim=4, stoch_discrete=8, hidden_dim=40, cnn_depth=4, action_dim=3,
                                batch_size=3, batch_length=4, imag_horizon=3, actor_grad="reinforce")),
}


def make_conf(preset="atari", **overrides):
    sections, over = PRESETS[preset]
    d = dict(DEFAULTS)
    for s in sections:
        d.update(SECTIONS[s])
    d.update(over)
    d.update(overrides)
    return Namespace(**d)
