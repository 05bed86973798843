from bayesian_torch_b200.utils.util import get_rho  # noqa: F401
