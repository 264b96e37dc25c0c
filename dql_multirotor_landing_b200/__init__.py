"""dql_multirotor_landing_b200: B200 (sm_100a) implementation of the batched landing-MDP step + tabular Double-Q
update of valerio98-lab/DQL_multirotor_landing, behind the reference's Python API.  See DESIGN.md."""
import os
from pathlib import Path

# PKG/__init__.py:5-7 resolves ASSETS_PATH through rospkg; here it is the repo's assets/ (or $DQL_ASSETS_PATH).
ASSETS_PATH = Path(os.environ.get("DQL_ASSETS_PATH", Path(__file__).resolve().parent.parent / "assets"))

__all__ = ["ASSETS_PATH"]
