"""python -m circkit_b200 canonicalize|uniq ... -- see circkit_b200/cli.py"""
import sys

from .cli import main

sys.exit(main())
