"""The reference LogWrappers' per-episode statistics written as array formulas (test helper).

classification: active_classification_env.py:149-197 -- per real step correct_label_prob = softmax(pred)[target];
at episode end accuracy = prob > 1/K, avg_/final_ of both, first_correct, last_incorrect.
regression: active_regression_env.py:204-227 -- euclidean_distance = ||target - pred||, mse = mean((target-pred)^2).
`preds`/`targets` are (T, N, ...) stacks over the episode's real steps."""
import numpy as np


def _softmax(x):
    e = np.exp(x - np.amax(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def classification_episode_stats(preds, targets, K):
    T, N = targets.shape
    prob = np.take_along_axis(_softmax(preds.astype(np.float32)), targets[..., None].astype(np.int64), -1)[..., 0]
    prob = prob.astype(np.float32)
    correct = prob > 1 / K
    acc = correct.astype(np.float32)
    first = np.where(correct.any(0), correct.argmax(0), -1)
    last_inc = np.where((~correct).any(0), T - 1 - (~correct)[::-1].argmax(0), -1)
    return {"avg_correct_label_prob": prob.mean(0, dtype=np.float32), "final_correct_label_prob": prob[-1],
            "avg_accuracy": acc.mean(0, dtype=np.float32), "final_accuracy": acc[-1],
            "first_correct": first.astype(np.int32), "last_incorrect": last_inc.astype(np.int32)}


def regression_episode_stats(preds, targets, K=0):
    d = targets.astype(np.float32) - preds.astype(np.float32)
    dist = np.sqrt(np.sum(d * d, axis=-1))
    mse = np.mean(d * d, axis=-1)
    return {"avg_euclidean_distance": dist.mean(0, dtype=np.float32), "final_euclidean_distance": dist[-1],
            "avg_mse": mse.mean(0, dtype=np.float32), "final_mse": mse[-1]}
