"""``simple_knn`` stand-in (the reference builds it from litegs/submodules/simple-knn; SURVEY 2.1 marks the extension itself
out of scope).  Only ``simple_knn._C.distCUDA2`` is used: litegs/scene/point.py:9."""
