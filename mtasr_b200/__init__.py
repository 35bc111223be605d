"""Importable alias of the product package.

The package directory is named after the reference repo (`multi-talker-asr-with-llms_b200/`), which is not a valid
Python identifier; `import mtasr_b200` resolves its sub-modules from that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "multi-talker-asr-with-llms_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
