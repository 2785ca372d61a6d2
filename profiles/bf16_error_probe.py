import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from oracle import net as onet
from oracle.games import make_game
from oracle.synth import random_positions, synthetic_state_dict
from test_net_gpu import device_net
game = make_game("othello", 8)
states = random_positions(game, 512, seed=2)
player = np.array([s[0] for s in states], np.uint8); first = np.array([s[1] for s in states], np.uint64); second = np.array([s[2] for s in states], np.uint64)
for cg, fs in [(0.577,0.577),(1.0,1.0),(1.414,1.0),(1.6,2.0)]:
    sd = synthetic_state_dict(game, 128, 3, 128, seed=21, conv_gain=cg, fc_scale=fs)
    wv, wp = onet.evaluate_batch(game, sd, states)
    net = device_net("othello", 8, 128, 3, 128, sd, "bf16", max_batch=1024)
    v, p = net.evaluate(player, first, second)
    rv, lg = net.forward(player, first, second)
    ov, ol = onet.forward(game, sd, onet.states_to_planes(game, states))
    print(f"gain {cg} fc {fs}: max|dv| {np.abs(v-wv).max():.4f} max|dp| {np.abs(p-wp).max():.4f} mean|dp| {np.abs(p-wp).mean():.5f} maxprob {wp.max():.3f} logit range {ol.numpy().min():.2f}..{ol.numpy().max():.2f} max|dlogit| {np.abs(lg-ol.numpy()).max():.4f}")
