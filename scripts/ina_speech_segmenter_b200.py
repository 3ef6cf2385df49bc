#!/usr/bin/env python
"""Drop-in for the reference's scripts/ina_speech_segmenter.py (same flags)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inaspeechsegmenter_b200.cli import main  # noqa: E402

if __name__ == '__main__':
    main()
