import sys; sys.path.insert(0, '.')
import numpy as np
from neumann_b200 import DeviceIndex
from neumann_b200.synth import synth_rows
idx = DeviceIndex(768); idx.fill_synthetic(2_000_000, 0x5EED0001)
q = synth_rows(16, 768, 0x5EED1001)
for _ in range(2): idx.search(q, 10, "cosine")
