"""Algorithm registry surface (mirrors omnisafe/algorithms/__init__.py:L69-85)."""
from omnisafe_b200.algorithms import on_policy, registry  # noqa: F401
from omnisafe_b200.algorithms.on_policy import (CPO, FOCOPS, PPO, RCPO, NaturalPG,  # noqa: F401
                                                PolicyGradient, PPOLag, TRPO, TRPOLag)

ALGORITHMS = {'on-policy': tuple(on_policy.ON_POLICY)}
ALGORITHM2TYPE = {algo: algo_type for algo_type, algos in ALGORITHMS.items() for algo in algos}
