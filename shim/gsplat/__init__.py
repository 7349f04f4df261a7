"""PYTHONPATH shim: `import gsplat` -> fusionsense_b200's sm_100a implementation."""
import fusionsense_b200 as _f

_f.install_gsplat_shim(force=True)
from fusionsense_b200.gsplat import *  # noqa: F401,F403,E402
from fusionsense_b200.gsplat import __version__  # noqa: F401,E402
