"""`ptsemseg.models` of the drop-in overlay: get_model() and the seven model classes of the B200 path, under the names
the reference's train.py / test.py / trainer.py import (train.py:176, test.py:93, trainer.py:18)."""
from multiagentperception_b200.models import (All_agents, LearnWhen2Com, LearnWho2Com, MIMO_All_agents,  # noqa: F401
                                              MIMOcom, MIMOcomWho, Single_agent, get_model)

__all__ = ["get_model", "Single_agent", "All_agents", "LearnWho2Com", "LearnWhen2Com", "MIMOcom", "MIMO_All_agents",
           "MIMOcomWho"]
