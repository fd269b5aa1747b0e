"""On-device counterparts of `anakin.artiboost` (the online-synthesis layer of the reference): CCV-space sampler,
view engine, grasp/object engines, pose generator, render provider.  Same class names and call signatures as the
reference modules of the same file names; everything heavy goes through the C-ABI in include/artiboost_b200.h."""
DUMMY = "dummy"  # CONST.DUMMY, anakin/utils/misc.py:70

from .view_engine import ViewEngine  # noqa: E402,F401
from .scrambler import (Scrambler, RandomScrambler, RandomScrambler2, RandomScrambler3, NaiveScrambler, NullScrambler,
                        AxisLayer)  # noqa: E402,F401
from .refiner import Refiner, NullRefine, HORefiner  # noqa: E402,F401
from .preprocessor import PreProcessorPoseGenerator  # noqa: E402,F401
from .ovg_set import OVGSet  # noqa: E402,F401
from .grasp_engine import GraspEngine  # noqa: E402,F401
from .object_engine import ObjEngine  # noqa: E402,F401
from .renderer import Renderer, PointLight, DirectionalLight, make_mesh  # noqa: E402,F401
from .render_infra import RendererProvider  # noqa: E402,F401
from .rendered_dataset import RenderedDataset  # noqa: E402,F401
