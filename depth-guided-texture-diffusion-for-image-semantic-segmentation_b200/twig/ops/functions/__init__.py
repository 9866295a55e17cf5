from .texture_diffusion_func import *  # noqa: F401,F403
from .texture_diffusion_func import MessagePassingFunction  # noqa: F401
