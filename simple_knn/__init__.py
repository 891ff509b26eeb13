"""Drop-in alias of the reference's `simple_knn` package (submodules/simple-knn): `from simple_knn._C import
distCUDA2` (scene/gaussian_model.py:21) resolves to the sm_100a implementation in eogs2_b200 when this
repository root is on sys.path ahead of the reference's compiled extension."""
