"""Prints the measured integer-ALU (LOP3) roofline of cuda:0 as one JSON line (gstim_measure_lop3_peak)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stim_b200

print(json.dumps(stim_b200.measure_lop3_peak(0)))
