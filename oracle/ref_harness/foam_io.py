"""TEST INFRASTRUCTURE ONLY - the golden-vector generator hands its cases to the unmodified reference as OpenFOAM binary
files; the format code lives in the package (adfvm_b200/foam_io.py)."""
from adfvm_b200.foam_io import write_polymesh, write_field  # noqa: F401
