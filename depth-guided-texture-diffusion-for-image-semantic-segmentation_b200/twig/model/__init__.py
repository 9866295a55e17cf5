from .texture_diffuser import *  # noqa: F401,F403
