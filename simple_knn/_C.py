"""`simple_knn._C` (submodules/simple-knn/ext.cpp:15-17): exports distCUDA2."""
from eogs2_b200.simple_knn import distCUDA2  # noqa: F401
