"""Golden vectors for the featuriser: outputs of the UNMODIFIED reference class (/root/reference/featurizer.py) on a fixed
set of strings.  Run in the build container: python tests/golden/make_golden_featurizer.py"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from featurizer import OneHotFeaturizer  # noqa: E402

CHARSET = [' ', '#', '(', ')', '+', '-', '/', '1', '2', '3', '4', '5', '6', '7', '8', '=', '@', 'B', 'C', 'F', 'H', 'I', 'N',
           'O', 'P', 'S', '[', '\\', ']', 'c', 'l', 'n', 'o', 'r', 's']           # 35 entries, as the ZINC charset of train.py:45-60
SMILES = ["CC(C)(C)c1ccc2occ(CC(=O)Nc3ccccc3F)c2c1", "C[C@@H]1CC(Nc2cncc(-c3nncn3C)c2)C[C@@H](C)C1", "N#Cc1ccc(-c2ccc(O[C@@H](C(=O)N3CCCC3)c3ccccc3)cc2)cc1",
          "c1ccccc1", "", "O", "[NH3+]CC(=O)[O-]", "Brc1ccc(/C=C\\c2ccccn2)cc1", "CCS(=O)(=O)N1CCC(C(=O)N2CCc3ccccc32)CC1"]
f = OneHotFeaturizer(CHARSET, 120)
onehot = f.featurize(SMILES)
decoded = f.one_hot_decode(onehot)
idx = onehot.argmax(-1)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "featurizer.npz"), charset=np.array(CHARSET), smiles=np.array(SMILES),
                    onehot=onehot.astype(np.uint8), ids=idx.astype(np.uint8), decoded=np.array([d[0] for d in decoded]),
                    from_index=np.array([f.decode_smiles_from_index(list(r)) for r in idx]))
print("wrote featurizer.npz", onehot.shape)
