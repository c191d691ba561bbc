"""Import-name shim so that ``from simple_knn._C import distCUDA2`` resolves to partgs_b200."""
