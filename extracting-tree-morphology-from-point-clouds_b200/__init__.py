"""B200-native nearest-cylinder label + offset path.

Import as ``treemorph_b200`` (the alias package at the repo root).  Submodules:

* ``synth``     synthetic QSMs / clouds
* ``build``     nvcc build of ``csrc/`` into ``libtreemorph_nn.so`` (the C-ABI library)
* ``binding``   ctypes binding of ``include/treemorph_nn.h``
* ``api``       device-pointer level Python API (torch tensors in, torch tensors out)
* ``sharding``  one-process-per-GPU point sharding with a single cylinder-table broadcast
* ``PreProcessing.LabelGenerationCuda`` / ``Modules.Projection``  drop-ins with the reference's signatures
"""
__version__ = "0.1.0"
