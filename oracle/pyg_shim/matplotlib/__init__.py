"""Import-only stub (test infrastructure): the reference imports `matplotlib.pyplot` at
datasets/PowerFlowData.py:13 without using it on the hot path; matplotlib is not installed here."""
