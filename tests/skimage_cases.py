"""scikit-image's documented behaviour of the three primitives the reference's sweep calls, as hand-written cases:
the docstring examples of ``skimage.measure.label`` and ``skimage.morphology.remove_small_objects`` (unchanged from
scikit-image 0.13 to 0.19, i.e. across the reference's 2018 time frame) and the border rules of
``skimage.morphology.binary_erosion`` / ``binary_dilation`` (binary.py: ``ndi.binary_erosion(image, structure=selem,
border_value=True)``, ``ndi.binary_dilation(image, structure=selem)``, default ``selem`` = the cross-shaped
connectivity-1 footprint).  scikit-image itself is not installed in this image; these cases are what pins the oracle
(tests/test_sweep_oracle.py) and, through it, the kernels (tests/test_gpu_sweep.py) to its semantics."""
import numpy as np

# skimage.measure.label docstring: x = np.eye(3); label(x, connectivity=2) (= the 2-D default) is ONE component,
# connectivity=1 would give three
LABEL_EYE = np.eye(3, dtype=bool)
LABEL_EYE_EXPECT = np.eye(3, dtype=np.int64)            # labels: 1 on the diagonal

# raster-order numbering: components are numbered by their first pixel in row-major order
LABEL_ORDER = np.array([[0, 1, 0, 0, 1],
                        [0, 0, 0, 0, 1],
                        [1, 0, 0, 0, 0],
                        [1, 1, 0, 1, 0]], dtype=bool)
LABEL_ORDER_EXPECT = np.array([[0, 1, 0, 0, 2],
                               [0, 0, 0, 0, 2],
                               [3, 0, 0, 0, 0],
                               [3, 3, 0, 4, 0]], dtype=np.int64)

# skimage.morphology.remove_small_objects docstring
RSO = np.array([[0, 0, 0, 1, 0],
                [1, 1, 1, 0, 0],
                [1, 1, 1, 0, 1]], dtype=bool)
RSO_MIN7_CONN2 = np.array([[0, 0, 0, 1, 0],            # remove_small_objects(a, 7, connectivity=2): the diagonal pixel belongs
                           [1, 1, 1, 0, 0],            # to the 7-pixel component, the lone pixel goes
                           [1, 1, 1, 0, 0]], dtype=bool)
RSO_MIN8_CONN2 = np.zeros((3, 5), dtype=bool)           # strictly-smaller-than rule: 7 < 8, everything goes

# erosion keeps set pixels on the border (pixels beyond the border count as set) ...
ERODE_FULL = np.ones((3, 4), dtype=bool)
ERODE_FULL_EXPECT = np.ones((3, 4), dtype=bool)
# ... and removes the cross around an unset pixel
ERODE_HOLE = np.ones((5, 5), dtype=bool)
ERODE_HOLE[2, 2] = False
ERODE_HOLE_EXPECT = np.ones((5, 5), dtype=bool)
ERODE_HOLE_EXPECT[2, 1:4] = False
ERODE_HOLE_EXPECT[1:4, 2] = False
# dilation of one pixel is the cross, clipped at the border (pixels beyond the border count as unset)
DILATE_DOT = np.zeros((4, 4), dtype=bool)
DILATE_DOT[0, 3] = True
DILATE_DOT_EXPECT = np.zeros((4, 4), dtype=bool)
DILATE_DOT_EXPECT[0, 2:4] = True
DILATE_DOT_EXPECT[1, 3] = True
# erosion then dilation (the reference's "get rid of singleton pixels", :149-152): a singleton and a 2-pixel bar vanish,
# a plus shape survives as itself, a 3 x 3 block survives as itself
OPEN_IN = np.zeros((9, 12), dtype=bool)
OPEN_IN[1, 1] = True
OPEN_IN[1, 4:6] = True
OPEN_IN[4, 2] = OPEN_IN[3, 2] = OPEN_IN[5, 2] = OPEN_IN[4, 1] = OPEN_IN[4, 3] = True
OPEN_IN[5:8, 7:10] = True
OPEN_EXPECT = OPEN_IN.copy()
OPEN_EXPECT[1, 1] = False
OPEN_EXPECT[1, 4:6] = False
OPEN_EXPECT[5, 7] = OPEN_EXPECT[5, 9] = OPEN_EXPECT[7, 7] = OPEN_EXPECT[7, 9] = False   # the block's corners: opening with a cross
