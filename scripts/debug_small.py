import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lamp_b200 as et
from tests.helpers import synth_classification, synth_regression
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
x, y = synth_classification(n, 6, 3, 1)
print("building cls", n, flush=True)
f = et.buildForestClassification(x, y, None, 3, 2, 3, 2, 2, seed=3)
print("cls ok", f.stats, flush=True)
x, y = synth_regression(n, 6, 1)
f = et.buildForestRegression(x, y, 2, 3, 2, 2, seed=3)
print("reg ok", f.stats, flush=True)
