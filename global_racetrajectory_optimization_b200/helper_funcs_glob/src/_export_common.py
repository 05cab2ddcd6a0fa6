"""Header shared by the two CSV exports: a random UUID and the SHA1 of the ggv file (or of an empty array's bytes)."""
import hashlib
import uuid

import numpy as np


def write_csv(path: str, ggv_path, table: np.ndarray, columns: tuple) -> None:
    if ggv_path is not None:
        with open(ggv_path, "rb") as fh:
            digest = hashlib.sha1(fh.read()).hexdigest()
    else:
        digest = hashlib.sha1(np.array([])).hexdigest()
    with open(path, "w") as fh:
        fh.write("# %s\n# %s\n" % (uuid.uuid4(), digest))
    with open(path, "ab") as fh:
        np.savetxt(fh, table, fmt="; ".join(["%.7f"] * len(columns)), header="; ".join(columns))
