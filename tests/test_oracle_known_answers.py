"""Pins the CPU oracle on the closed-form known answers of the reference's own aggregate / update / delete tests
(SURVEY.md 8c G4, G6); the cases live in tests/known_answer_cases.py and also run through CUDA (test_gpu_known_answers.py)."""
import pytest

import known_answer_cases as K


@pytest.mark.parametrize("case", K.CASES, ids=lambda c: c.__name__)
def test_known_answer_on_oracle(case, oracle_api):
    case(K.OracleEngine(oracle_api))
