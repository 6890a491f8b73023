"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/otsu_cases.npz: the Otsu thresholds and 256-bin histograms of the seeded
predictions of ``port_norm.otsu_cases()``.  scikit-image is not available in this image (see ``port_norm.threshold_otsu``), so
the fixture is produced by the restatement on top of numpy's own ``np.histogram``; it pins that behaviour across numpy versions.

    python -m oracle.make_golden_otsu
"""
import os

import numpy as np

from . import port_norm


def main():
    out = {}
    for name, img in port_norm.otsu_cases().items():
        out["th." + name] = np.asarray(port_norm.threshold_otsu(img))
        if not np.all(img == img.reshape(-1)[0]):
            out["counts." + name] = np.histogram(img.reshape(-1), bins=256)[0]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "otsu_cases.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.tolist() if v.ndim == 0 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
