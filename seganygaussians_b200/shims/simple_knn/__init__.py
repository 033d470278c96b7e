"""Stand-in for the ``simple_knn`` extension (``submodules/simple-knn``), backed by libsagars' exact grid KNN.
The reference imports ``from simple_knn._C import distCUDA2`` (scene/gaussian_model.py:20, gaussian_model_ff.py:21)."""
