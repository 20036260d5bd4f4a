"""Configuration blocks accepted by daydreamer_b200.Agent.

`Agent.configs` plays the role of the reference's configs.yaml
(embodied/agents/dreamerv2plus/configs.yaml, read at agent.py:18-19): the
reference's train.py builds `embodied.Config(Agent.configs['defaults'])`, applies
named blocks, then command-line overrides.  Keys and default values follow the
reference so its flags (`--configs a1 --rssm.deter 512 ...`) keep working; keys
that only configured TensorFlow (`tf.*`) are accepted and ignored.  Blocks that
only select environments the learner never touches are not restated here.
"""

_dense = dict(act='elu', norm='layer')
_opt = dict(opt='adam', lr=1e-4, eps=1e-6, clip=100.0, wd=1e-2, wd_pattern='kernel', warmup=0)
_head = dict(layers=4, units=512, outscale=0.1, inputs=['deter', 'stoch'], **_dense)
_norm = lambda impl, decay: dict(impl=impl, decay=decay, max=1e8)

DEFAULTS = dict(
    # run loop (consumed by embodied/agents/dreamerv2plus/train.py and embodied.run.*)
    logdir='/dev/null', run='train_fixed_eval', seed=0, task='dummy_discrete',
    env=dict(amount=4, parallel='process', daemon=False, repeat=1, size=[64, 64], camera=-1,
             gray=False, length=0, discretize=0, lives=False, sticky=True, episodic=True,
             restart=True, resets=True, seed=0, kbreset=False),
    replay='fixed', replay_size=1e6, replay_chunk=64,
    replay_fixed=dict(length=0, prio_starts=0.0, prio_ends=1.0, sync=0, minlen=0),
    replay_consec=dict(randomize=False, sync=0),
    replay_prio=dict(prio_starts=0.0, prio_ends=1.0, sync=0, fraction=0.1, softmax=False,
                     temp=1.0, constant=0.0, exponent=0.5),
    tf=dict(jit=True, platform='gpu', precision='float16', debug_nans=False, logical_gpus=0,
            dist_dataset=False, dist_policy=False, tensorfloat=True, soft_placement=False,
            growth=True),
    eval_dir='', filter='.*', tbtt=0,
    train=dict(steps=1e10, expl_until=0, log_every=1e4, eval_every=1e5, eval_eps=1,
               eval_samples=1, train_every=16, train_steps=1, train_fill=1e4, eval_fill=1e4,
               pretrain=1, log_zeros=False, log_keys_video=['image'], log_keys_sum='^$',
               log_keys_mean='^$', log_keys_max='^$', sync_every=180),
    # agent
    task_behavior='Greedy', expl_behavior='None', transform_rewards='off', batch_size=32,
    data_loader='embodied', expl_noise=0.0, eval_noise=0.0, priority='reward_loss',
    priority_correct=0.0, train_wm=True, known_reward='none',
    # world model
    grad_heads=['decoder', 'reward', 'cont'],
    rssm=dict(units=1024, deter=1024, stoch=32, classes=32, initial='learned2', unimix=0.01,
              prior_layers=3, post_layers=1, gru_layers=1, unroll=True, **_dense),
    encoder=dict(mlp_keys='.*', cnn_keys='.*', mlp_layers=4, mlp_units=512, cnn='simple',
                 cnn_depth=64, cnn_kernels=[4, 4, 4, 4], cnn_blocks=2, **_dense),
    decoder=dict(mlp_keys='.*', cnn_keys='.*', mlp_layers=4, mlp_units=512, cnn='simple',
                 cnn_depth=64, cnn_kernels=[5, 5, 6, 6], cnn_blocks=2, image_dist='mse',
                 inputs=['deter', 'stoch'], **_dense),
    reward_head=dict(dist='symlog', **_head),
    cont_head=dict(dist='binary', **_head),
    loss_scales=dict(kl=1.0, image=1.0, reward=1.0, cont=1.0),
    model_opt=dict(_opt),
    wmkl=dict(impl='mult', scale=0.1, target=3.5, min=1e-3, max=1.0, vel=0.1),
    wmkl_balance=0.8,
    # actor critic
    actor=dict(minstd=0.03, maxstd=1.0, unimix=0.01, **_head),
    critic=dict(dist='symlog', **_head),
    actor_opt=dict(_opt), critic_opt=dict(_opt),
    actor_dist_disc='onehot', actor_dist_cont='normal', discount=0.998, imag_horizon=15,
    imag_unroll=True, return_lambda=0.95, actor_return='gve', critic_return='gve',
    actor_grad_disc='reinforce', actor_grad_cont='backprop', slow_target=True,
    slow_target_update=100, slow_target_fraction=1.0,
    actent=dict(impl='mult', scale=3e-3, target=0.5, min=1e-5, max=1.0, vel=0.1),
    actent_norm=True, advnorm=_norm('mean_std', 0.99), retnorm=_norm('off', 0.999),
    scorenorm=_norm('off', 0.99), critic_type='vfunction',
    # exploration (only `expl_behavior: None | Random` is built; the keys of A/configs.yaml:91-112 are
    # kept so that reference command lines which set them still parse)
    expl_rewards=dict(extr=0.0, disag=0.0, vae=0.0, ctrl=0.0, pbe=0.0),
    expl_discount=0.99, expl_retnorm=_norm('std', 0.999), expl_scorenorm=_norm('off', 0.999),
    expl_opt=dict(opt='adam', lr=1e-4, eps=1e-6, clip=100.0, wd=1e-2, warmup=0),
    disag_head=dict(layers=4, units=512, act='elu', norm='layer', dist='mse', outscale=1.0,
                    inputs=['deter', 'stoch', 'action']),
    disag_target=['stoch'], disag_models=8,
    ctrl_embed=dict(layers=3, units=512, act='elu', norm='layer', dist='mse', inputs=['deter', 'stoch']),
    ctrl_head=dict(layers=1, units=128, act='elu', norm='layer', dist='mse', inputs=['current', 'next']),
    ctrl_size=32, ctrl_opt=dict(opt='adam', lr=1e-4, eps=1e-6, clip=100.0, wd=1e-2, warmup=0),
    pbe_inputs=['deter'], pbe_knn=16,
    expl_enc=dict(layers=4, units=512, act='elu', norm='layer', dist='onehot', outscale=0.1,
                  inputs=['deter'], shape=[8, 8]),
    expl_dec=dict(layers=4, units=512, act='elu', norm='layer', dist='mse', outscale=0.1),
    expl_kl=dict(impl='mult', scale=0.1, target=10.0, min=1e-2, max=1.0, vel=0.1),
    expl_vae_elbo=False, expl_when_buffer=1e4, expl_when_frac=0.1, expl_when_every=50,
    expl_when_random=False,
)

_robot = {
    'run': 'train', 'batch_size': 32, 'replay_chunk': 32, 'env.amount': 1,
    'env.parallel': 'none', 'env.restart': False, 'train.train_every': 4,
    'discount': 0.995, 'train.sync_every': 20, 'actor.minstd': 0.1, 'actor.unimix': 0.1,
}

CONFIGS = {
    'defaults': DEFAULTS,
    'a1': {
        **_robot, 'task': 'a1_sim',
        'encoder': {'mlp_keys': 'vector', 'cnn_keys': '$^'},
        'decoder': {'mlp_keys': 'vector', 'cnn_keys': '$^'},
        'env.repeat': 50, 'env.resets': False, 'env.length': 250,
        'actent.target': 0.5, 'rssm': {'deter': 256, 'units': 256},
        'train.train_fill': 5000,
    },
    'xarm': {
        **_robot, 'task': 'xarm_dummy',
        'encoder': {'mlp_keys': 'cartesian|joint|gripper|grasped', 'cnn_keys': 'image|depth'},
        'decoder': {'mlp_keys': 'cartesian|joint|gripper|grasped', 'cnn_keys': 'image|depth'},
        'env.repeat': 1, 'actent.target': 0.4, 'rssm': {'deter': 512, 'units': 512},
        'train.log_every': 100, 'train.eval_every': 20,
        'train.log_keys_video': ['image', 'depth'], 'env.resets': True, 'env.length': 100,
        'train.train_fill': 2000,
    },
    'ur5': {
        **_robot, 'task': 'ur5_dummy',
        'encoder': {'mlp_keys': 'cartesian|joint|gripper|grasped', 'cnn_keys': 'image'},
        'decoder': {'mlp_keys': 'cartesian|joint|gripper|grasped', 'cnn_keys': 'image'},
        'env.repeat': 1, 'actent.target': 0.4, 'rssm': {'deter': 512, 'units': 512},
        'train.log_every': 100, 'train.eval_every': 20,
        'train.log_keys_video': ['image', 'depth'], 'env.resets': True, 'env.length': 100,
        'train.train_fill': 2000,
    },
    'dmc_proprio': {
        'task': 'dmc_walker_walk',
        'encoder': {'mlp_keys': '.*', 'cnn_keys': '$^'},
        'decoder': {'mlp_keys': '.*', 'cnn_keys': '$^'},
        'env.repeat': 2, 'train.train_every': 2,
    },
    'small': {
        'env.amount': 2,
        'train': {'eval_every': 1e4, 'log_every': 1e3, 'train_fill': 100, 'eval_fill': 100,
                  'train_every': 16},
        'tf.precision': 'float32', 'replay_chunk': 16, 'rssm.deter': 128, 'rssm.units': 128,
        'rssm.stoch': 8, r'.*\.cnn_depth$': 16, r'.*\.wd$': 0.0,
    },
    'debug': {
        'env.length': 100, 'env.restart': False, 'env.amount': 2,
        'train': {'eval_every': 300, 'log_every': 300, 'train_fill': 100, 'eval_fill': 100,
                  'train_steps': 1, 'train_every': 30},
        'batch_size': 8, 'replay_size': 500, 'replay_chunk': 12,
        'encoder.cnn_depth': 16, 'decoder.cnn_depth': 16,
        'rssm': {'units': 64, 'stoch': 8, 'classes': 8},
        r'.*\.layers': 2, r'.*\.units': 64, r'.*\.wd': 0.0,
        'tf': {'platform': 'gpu', 'jit': False},
    },
}
