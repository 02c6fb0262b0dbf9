"""CPU ORACLE / BASELINE (test infrastructure, NOT product code): the reference's per-frame *numpy* path, restated with
the same numpy calls so that its cost per frame is the reference's cost per frame.

The reference itself cannot travel to the GPU box (/root/reference does not exist there) and its main.py cannot be
imported without gym-retro/DEAP, so bench.py's second CPU baseline ("port-numpy") runs this restatement: frames come from
the C oracle emulator (oracle.Atari, standing in for gym-retro's env.step) and everything above the emulator is numpy /
Python as in the reference:

  find_stuff / get_rect_quickly        utils.py:14-19, 60-68   (full-frame compare + argwhere + average per colour)
  NeuralNetwork.populate / run         numpy_nn.py:52-69, 120-137
  inference                            utils.py:139-153
  keep_within_game_bounds_please       utils.py:71-77
  get_actions                          main.py:138-154
  calculate_timeout_and_frames         main.py:128-135
  perform_episode                      main.py:69-112
  HardcodedAi                          dumb_ais.py:1-8

tests/test_numpy_path.py pins it to the goldens generated from the imported reference (tests/golden/*.npz)."""
from __future__ import annotations

import warnings

import numpy as np

TOP, BOTTOM, WIDTH, PLAYABLE = 34, 194, 160, 160        # config.py:8-13
PADDLE_H = 16.0                                           # config.py:10
COLOURS = ((236, 236, 236), (213, 130, 74), (92, 186, 92))     # ball, left, right: config.py:4-6


def get_rect_quickly(chopped, colour):
    """Mean (row, col) over every per-channel match, None when nothing matches (utils.py:60-68)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            centre = np.average(np.argwhere(chopped == colour)[:, :-1], axis=0)
    return None if np.isnan(centre).any() else centre


def find_stuff(observation):
    crop = observation[TOP:BOTTOM, :]
    return [get_rect_quickly(crop, c) for c in COLOURS]


def sigmoid(x):
    return 1 / (1 + np.e ** -x)          # numpy_nn.py:22-23


class NeuralNetwork:
    def __init__(self, nodes, weights, bias=True):
        self.nodes = list(nodes)
        self.extra = 1 if bias else 0
        self.cut = -1 if bias else None
        self.layers = []
        self.buffers = [np.ones(n + self.extra) for n in self.nodes]       # trailing 1.0 = bias input
        used = 0
        w = np.asarray(weights, np.float64)
        for a, b in zip(self.nodes[:-1], self.nodes[1:]):
            count = (a + self.extra) * b
            self.layers.append(w[used:used + count].reshape(b, a + self.extra))      # bias weight = last column
            used += count

    def run(self, vec):
        if len(vec) != self.nodes[0]:
            raise Exception("input vector wrong shape")
        self.buffers[0][:self.cut] = vec
        for i, m in enumerate(self.layers):
            self.buffers[i + 1][:self.cut] = sigmoid(np.dot(m, self.buffers[i]))
        best = np.argmax(self.buffers[-1][:self.cut])
        return [1, 0] if best == 0 else [0, 1]


class HardcodedAi:
    def run(self, v):
        if v[1] < v[4]:
            return [1, 0]
        if v[1] > v[4]:
            return [0, 1]
        return [0, 0]


def inference(ball, last, me, enemy, model):
    return model.run([ball[1] / WIDTH, ball[0] / PLAYABLE, last[1] / WIDTH, last[0] / PLAYABLE, me[0] / PLAYABLE, enemy[0] / PLAYABLE])


def keep_within_bounds(paddle, action):
    if paddle is not None:
        if paddle[0] < PADDLE_H:
            return [0, 1]
        if paddle[0] > (BOTTOM - TOP) - PADDLE_H:
            return [1, 0]
    return action


def get_actions(ball, last, left, left_model, right, right_model, rnd):
    eye = np.eye(2, dtype=int)
    la, ra = eye[rnd(), :], eye[rnd(), :]                # two draws every frame (main.py:139-140)
    if last is None:
        last = ball
    if ball is None:
        return [0, 0], [0, 0]
    if left is not None and right is not None:           # (the reference raises TypeError when only one paddle is visible)
        la = inference([ball[0], WIDTH - ball[1]], [last[0], WIDTH - last[1]], left, right, left_model)
        ra = inference(ball, last, right, left, right_model)
    return la, ra


def perform_episode(emu, left_model, right_model, mult=1.0, rnd=None, max_frames=0, win_score=3, timeout_thresh=2000):
    """One game on an oracle.Atari that has been reset to its start state.  Returns (env.step calls, reward)."""
    import oracle
    rnd = rnd or (lambda: np.random.choice(2))
    action = np.zeros(16, dtype=int)
    action[0] = 1; action[-1] = 1                        # BLANK_ACTION
    last_score, timeout, total, last_ball, steps = None, 0.0, 0.0, None, 0
    while True:
        fb = emu.step(action)
        obs = oracle.fb_to_rgb(fb)
        ram = emu.ram
        score = {"score1": int(ram[13]), "score2": int(ram[14])}
        steps += 1
        ball, left, right = find_stuff(obs)
        la, ra = get_actions(ball, last_ball, left, left_model, right, right_model, rnd)
        last_ball = ball
        action[4:6] = keep_within_bounds(right, ra)
        action[6:8] = keep_within_bounds(left, la)
        if last_score is not None:
            if last_score == score:
                timeout += 1.0
            else:
                total += timeout; timeout = 0.0
        last_score = score
        if score["score1"] >= win_score or score["score2"] >= win_score or timeout > timeout_thresh:
            break
        if max_frames and steps >= max_frames:
            break
    if score["score1"] == score["score2"]:
        return steps, 0.0
    diff = score["score2"] - score["score1"]
    return steps, (diff + score["score2"] * mult) / (total / 100.0)
