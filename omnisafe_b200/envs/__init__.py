"""Environment side of the hot path: the HBM-resident synthetic Box CMDP."""
from omnisafe_b200.envs.synthetic import SyntheticBoxEnv, env_bias, support_envs  # noqa: F401
