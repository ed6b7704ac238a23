"""Development tools only: PYJAC_B200_LIB=path selects another build of the library (tools/devbuild.sh) for
the tool that imports this module.  The product itself reads no environment variable."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.environ.get('PYJAC_B200_LIB'):
    from pyjac_b200 import lib
    lib.use(os.environ['PYJAC_B200_LIB'])
