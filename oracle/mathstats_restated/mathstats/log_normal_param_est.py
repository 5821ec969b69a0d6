"""ORACLE / TEST INFRASTRUCTURE -- placeholder for `mathstats.log_normal_param_est`.
The lognormal scoring branch of the reference (CreateGraph.py:485-493,523-531)
raises TypeError under Python 3 (`range` with a float step, :490), so it is a
"next" row and is not restated."""


def GapEstimator(mu, sigma, read_length, samples, c1_len, c2_len=None):
    raise NotImplementedError("lognormal GapEstimator is out of scope (SURVEY.md 8f rank 3)")
