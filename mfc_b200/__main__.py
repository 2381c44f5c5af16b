"""``python -m mfc_b200 run case.py`` -- see ``mfc_b200.main``."""
import sys

from .main import main

sys.exit(main())
