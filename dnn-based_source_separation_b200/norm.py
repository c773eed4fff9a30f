from modules.norm import *  # noqa: F401,F403  (deprecation shim like the reference's src/norm.py:3)
