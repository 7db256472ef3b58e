"""Golden vectors for the learners' update rule -- TEST INFRASTRUCTURE.

Keras / TensorFlow are not installable here, so the reference's `replay` cannot be run; this script restates, in plain
NumPy float64 and WITHOUT autograd, exactly what one `replay()` of each reference learner computes, from the reference's
source and from the Keras 2 code paths it calls, and writes the result as tests/golden/agents/n1_replay.npz:

  targets      DQN.replay, DQN.py:156-185 (max bootstrap) / DQN_SARSA.replay, DQN_SARSA.py:103-132 (Q(s', a')):
               per sample `prediction = target.predict(state)[0]; prediction[action] = reward if done else
               reward + gamma * bootstrap` -- a Python loop over the batch, as the reference writes it.
  networks     DQN.make_network, DQN.py:209-233: Flatten -> Dense(50, sigmoid) -> Dense(A, linear);
               DQN_DUEL.make_network, DQN_DUEL.py:18-48: two such streams, q = value + (advantage - mean(advantage)).
  fit          `model.fit(states, predictions, epochs=1)` with batch_size 32 = the Keras default mini-batch: ONE update.
               loss 'mse' = mean over outputs, then over the batch (keras.losses.mean_squared_error + sample mean).
  optimizer    Adam(lr=alpha, clipvalue=1), Keras 2 `Adam.get_updates`: every gradient element clipped to [-1, 1];
               lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t); m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
               p -= lr_t * m / (sqrt(v) + 1e-7)   (epsilon = K.epsilon(), beta_1 = 0.9, beta_2 = 0.999).
  two updates  per learner (t = 1, 2), so that the moment estimates carry over.  Stored: the weights before (float32, what
               the learner under test is loaded with), every batch, its targets and loss, and the weights after each
               update (second layers whole, of the 300-row first-layer kernels every 16th row).

Gradients are written out by hand (chain rule of the two-layer streams), so the fixture does not depend on the autograd
it is used to check.  Run:  python oracle/gen_n1_fixture.py
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GAMMA, ALPHA, B1, B2, EPS = 0.999, 0.005, 0.9, 0.999, 1e-7  # constants.py:53-54; keras.optimizers.Adam defaults
N_IN, HID, A, BATCH = 10 * 10 * 3, 50, 4, 32


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def forward(w, x, duel):
    """Q-values [B, A] and the activations the backward pass needs."""
    h1 = sigmoid(x @ w["dense_1/kernel:0"] + w["dense_1/bias:0"])
    adv = h1 @ w["dense_2/kernel:0"] + w["dense_2/bias:0"]
    if not duel:
        return adv, (h1,)
    h3 = sigmoid(x @ w["dense_3/kernel:0"] + w["dense_3/bias:0"])
    val = h3 @ w["dense_4/kernel:0"] + w["dense_4/bias:0"]
    return val + (adv - adv.mean(axis=1, keepdims=True)), (h1, h3)


def gradients(w, x, y, duel):
    """d(mse)/d(weights), by hand.  mse = mean over all B * A elements of (q - y)^2."""
    q, act = forward(w, x, duel)
    dq = 2.0 * (q - y) / q.size
    g = {}
    dadv = dq - dq.mean(axis=1, keepdims=True) if duel else dq  # d(adv_j - mean(adv)) / d adv_i = delta_ij - 1/A
    h1 = act[0]
    g["dense_2/kernel:0"] = h1.T @ dadv
    g["dense_2/bias:0"] = dadv.sum(axis=0)
    dz1 = (dadv @ w["dense_2/kernel:0"].T) * h1 * (1.0 - h1)
    g["dense_1/kernel:0"] = x.T @ dz1
    g["dense_1/bias:0"] = dz1.sum(axis=0)
    if duel:
        h3 = act[1]
        dval = dq.sum(axis=1, keepdims=True)
        g["dense_4/kernel:0"] = h3.T @ dval
        g["dense_4/bias:0"] = dval.sum(axis=0)
        dz3 = (dval @ w["dense_4/kernel:0"].T) * h3 * (1.0 - h3)
        g["dense_3/kernel:0"] = x.T @ dz3
        g["dense_3/bias:0"] = dz3.sum(axis=0)
    return g, float(((q - y) ** 2).mean())


def keras_adam_step(w, g, m, v, t):
    lr_t = ALPHA * np.sqrt(1.0 - B2 ** t) / (1.0 - B1 ** t)
    out = {}
    for k in w:
        gk = np.clip(g[k], -1.0, 1.0)  # clipvalue=1
        m[k] = B1 * m[k] + (1.0 - B1) * gk
        v[k] = B2 * v[k] + (1.0 - B2) * gk * gk
        out[k] = w[k] - lr_t * m[k] / (np.sqrt(v[k]) + EPS)
    return out


def replay_targets(target_w, batch, duel, sarsa):
    """The reference's loop, sample by sample."""
    s, a, r, sp, ap, d = batch
    out = []
    for i in range(len(a)):
        prediction = forward(target_w, s[i:i + 1], duel)[0][0].copy()
        if d[i]:
            prediction[a[i]] = r[i]
        else:
            q_next = forward(target_w, sp[i:i + 1], duel)[0][0]
            predQ = q_next[ap[i]] if sarsa else np.amax(q_next)
            prediction[a[i]] = r[i] + GAMMA * predQ
        out.append(prediction)
    return np.array(out)


def glorot(rng, n_in, n_out):
    lim = np.sqrt(6.0 / (n_in + n_out))
    return rng.uniform(-lim, lim, size=(n_in, n_out))


def make_weights(rng, duel):
    w = {"dense_1/kernel:0": glorot(rng, N_IN, HID), "dense_1/bias:0": rng.normal(0, 0.05, HID),
         "dense_2/kernel:0": glorot(rng, HID, A), "dense_2/bias:0": rng.normal(0, 0.05, A)}
    if duel:
        w.update({"dense_3/kernel:0": glorot(rng, N_IN, HID), "dense_3/bias:0": rng.normal(0, 0.05, HID),
                  "dense_4/kernel:0": glorot(rng, HID, 1), "dense_4/bias:0": rng.normal(0, 0.05, 1)})
    return w


def make_batch(rng):
    """32 transitions shaped like the environment's: 0/1 observations, rewards -1 / -1000 / +1000 / a burn-out fraction."""
    s = (rng.random((BATCH, N_IN)) < 0.35).astype(np.float64)
    sp = s.copy()
    flip = rng.random((BATCH, N_IN)) < 0.02
    sp[flip] = 1.0 - sp[flip]
    a = rng.integers(0, A, BATCH)
    ap = rng.integers(0, A, BATCH)
    kind = rng.integers(0, 8, BATCH)
    r = np.where(kind == 0, -1000.0, np.where(kind == 1, 1000.0, np.where(kind == 2, 1000.0 * 57 / 100, -1.0)))
    d = (kind == 0) | (kind == 2)
    return s, a, r, sp, ap, d


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    for name, duel, sarsa in (("DQN", False, False), ("DQN_SARSA", False, True), ("DQN_DUEL", True, False), ("DQN_BOTH", True, True)):
        w = make_weights(rng, duel)
        tw = {k: v + rng.normal(0, 0.03, v.shape) for k, v in w.items()}  # target != model: the wrong network would show
        m = {k: np.zeros_like(v) for k, v in w.items()}
        v = {k: np.zeros_like(x) for k, x in w.items()}
        for k, x in w.items():
            out[f"{name}/w0/{k}"] = x.astype(np.float32)
            out[f"{name}/target/{k}"] = tw[k].astype(np.float32)
        for t in (1, 2):
            batch = make_batch(rng)
            y = replay_targets(tw, batch, duel, sarsa)
            g, loss = gradients(w, batch[0], y, duel)
            w = keras_adam_step(w, g, m, v, t)
            for key, arr in zip(("s", "a", "r", "sp", "ap", "d"), batch):
                out[f"{name}/t{t}/batch_{key}"] = np.asarray(arr)
            out[f"{name}/t{t}/targets"] = y
            out[f"{name}/t{t}/loss"] = np.float64(loss)
            for k, x in w.items():  # the big first-layer kernels: every 16th input row (the fixture stays small)
                out[f"{name}/t{t}/w/{k}"] = (x[::16] if x.shape[0] == N_IN else x).astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "agents", "n1_replay.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
