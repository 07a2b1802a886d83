"""The host-side decision models against known answers produced by the reference's own learned_model.py
(tests/golden/make_learned_model_golden.py) — q-values bit-exact, discrete decisions identical."""
import golden_io
from interactive_highway_simulation_b200 import learned_model as lm


def test_against_reference_known_answers():
    (fixture,) = golden_io.load_all("learned_model")
    lc, mm = lm.LearnedLaneChangeModel(), lm.LearnedMergingModel()
    n = 0
    for c in fixture["cases"]:
        if c["model"] == "lane_change":
            q, d = lc.inference(c["features"], c["commands"], c["last"])
            assert d == c["decision"]
            assert [[k, float(v)] for k, v in q] == c["q"]
        else:
            idx, q = mm.inference(c["features"])
            assert idx == c["index"]
            assert {str(k): float(v) for k, v in q.items()} == c["q"]
        n += 1
    assert n >= 80


def test_survey_known_answer():
    """first case of the reference's __main__ -> `left`, q(left) = 1.5034182666958098."""
    (fixture,) = golden_io.load_all("learned_model")
    c = fixture["cases"][0]
    q, d = lm.LearnedLaneChangeModel().inference(c["features"], c["commands"])
    assert d == "left" and q[0][0] == "left" and float(q[0][1]) == 1.5034182666958098


def test_feature_vector_order():
    class F:
        successRate, fallBackRate, utility, comfort, expectedStepRatio, utilityObjects, comfortObjects = 1, 2, 3, 4, 5, 6, 7
    assert lm.feature_vector(F) == [1, 2, 3, 4, 5, 6, 7]
