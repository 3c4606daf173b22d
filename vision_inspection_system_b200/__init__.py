"""Import alias for ``vision-inspection-system_b200/`` (a hyphen cannot appear in an import statement).

``import vision_inspection_system_b200`` resolves every submodule from the hyphenated directory beside this one.
"""
import pathlib as _pathlib

_real = _pathlib.Path(__file__).resolve().parent.parent / "vision-inspection-system_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
