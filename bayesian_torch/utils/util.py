from bayesian_torch_b200.utils.util import MOPED, entropy, get_rho, mutual_information, predictive_entropy  # noqa: F401
