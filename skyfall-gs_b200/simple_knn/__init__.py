"""simple_knn — drop-in for the reference package (scene/gaussian_model.py:25 does `from simple_knn._C import distCUDA2`)."""
