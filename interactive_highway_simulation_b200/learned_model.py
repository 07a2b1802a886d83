"""The learned decision models on top of the rollout features (host side, binary64).

Mirrors the reference's learned_model.py (class names, attribute names, `inference` signatures and return shapes):
the "model" is a 7-weight dot product per candidate plus a threshold filter (merging / exit, learned_model.py:4-39) or
per-command costs with a hysteresis towards the previous decision (lane change, learned_model.py:42-73) — not a neural
network, so it stays on the host; the dot product is numpy's, as in the reference, so q-values agree bit for bit.
"""
import numpy as np

FEATURE_ORDER = ("successRate", "fallBackRate", "utility", "comfort", "expectedStepRatio", "utilityObjects", "comfortObjects")


def feature_vector(behavior_feature):
    """BehaviorFeature -> the list the reference builds at behaviors.py:106-107,127-128,166-167."""
    return [getattr(behavior_feature, name) for name in FEATURE_ORDER]


def filter_actions(valid_features, q_values, sr_threshold, fallback_threshold):
    """Best-q candidate whose success rate >= sr_threshold and fall-back rate <= fallback_threshold, else -1
    ("go after the last vehicle")."""
    if all(f[0] < sr_threshold for f in valid_features.values()):
        return -1
    ranked = sorted(q_values.items(), key=lambda kv: kv[1], reverse=True)
    for action, _ in ranked:
        f = valid_features[action]
        if f[0] >= sr_threshold and f[1] <= fallback_threshold:
            return action
    return -1


class ClonedBehaviorParameter:
    def __init__(self, min_success_rate, max_fall_back_rate):
        self.min_success_rate = min_success_rate
        self.max_fall_back_rate = max_fall_back_rate


class LearnedMergingModel:
    def __init__(self):
        self.weights = np.array([0.5, -0.7, 0.05, 0.05, -1, 0.1, 0.15])
        self.learned_model_parameter = ClonedBehaviorParameter(0.15, 1.)
        self.feature_placeholder = [0., 1., 0., 0., 1., 0., 0.]

    def inference(self, feature):
        """feature: one 7-vector per candidate gap -> (index of the chosen gap or -1, {index: q})."""
        valid = {i: f for i, f in enumerate(feature)}
        q_values = {i: np.dot(self.weights, np.array(f)) for i, f in enumerate(feature)}
        p = self.learned_model_parameter
        return filter_actions(valid, q_values, p.min_success_rate, p.max_fall_back_rate), q_values


class LearnedLaneChangeModel:
    def __init__(self):
        self.weights = np.array([0.183, -0.367, 0.3, 0.1, -0.15, 1.0, 0.25])
        self.feature_placeholder = [0., 1., 0., 0., 1., 0., 0.]
        self.left_cost = -0.01
        self.right_cost = 0.03
        self.delta_to_last_decision = 0.015

    def inference(self, feature, commands, last_decision="initial"):
        """-> ([[command, q], ...] sorted by q descending, decision); a new command must beat the previous decision's q
        by delta_to_last_decision."""
        if len(feature) != len(commands):
            print("Dimension of features do not match possible commands!")
        side_cost = {"left": self.left_cost, "right": self.right_cost}
        q_values = []
        for command, f in zip(commands, feature):
            q = np.dot(self.weights, np.array(f))
            if command in side_cost:
                q += side_cost[command]
            q_values.append([command, q])
        bar = dict(q_values).get(last_decision, -1e10) + self.delta_to_last_decision   # what a new command has to reach
        q_values.sort(key=lambda cq: cq[1], reverse=True)
        if q_values[0][0] == last_decision:
            return q_values, last_decision
        for command, q in q_values:
            if q >= bar:
                return q_values, command
        return q_values, last_decision
