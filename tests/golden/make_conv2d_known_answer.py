"""Writes tests/golden/conv2d_known_answer.json.

The reference is Nim and cannot be executed in this image, so the vectors are TRANSCRIBED from
the reference's own convolution self-check (`conv_impl_check`,
/root/reference/benchmarks/convolution/conv2d_common.nim:128-283), which every convolution
implementation of the reference (direct, im2col, MEC) must satisfy with `doAssert target == output`.
Layouts: input NCHW, kernel (c_out, c_in, kH, kW), output NCHW.  All values are small integers,
exact in fp32/tf32/bf16, so every path must reproduce them exactly.
Run:  python tests/golden/make_conv2d_known_answer.py
"""
import json
import os

CASES = [
    dict(src="conv2d_common.nim:137-178", ishape=[1, 1, 4, 4], kshape=[1, 1, 3, 3], padding=[1, 1], strides=[1, 1],
         input=[[[[1, 2, 0, 0], [5, 3, 0, 4], [0, 0, 0, 7], [9, 3, 0, 0]]]],
         kernel=[[[[1, 1, 1], [1, 1, 0], [1, 0, 0]]]],
         target=[[[[1, 8, 5, 0], [8, 11, 5, 4], [8, 17, 10, 11], [9, 12, 10, 7]]]]),
    dict(src="conv2d_common.nim:180-283", ishape=[1, 3, 5, 5], kshape=[2, 3, 3, 3], padding=[1, 1], strides=[2, 2],
         input=[[[[2, 2, 0, 2, 1], [0, 1, 1, 0, 2], [1, 2, 1, 2, 1], [2, 2, 0, 0, 2], [2, 1, 1, 1, 2]],
                 [[2, 0, 1, 1, 1], [2, 2, 0, 0, 2], [2, 2, 1, 0, 0], [1, 1, 2, 2, 0], [2, 1, 1, 1, 0]],
                 [[0, 1, 2, 2, 0], [1, 1, 1, 1, 0], [2, 1, 2, 2, 0], [0, 2, 2, 2, 1], [0, 0, 2, 2, 1]]]],
         kernel=[[[[-1, -1, -1], [1, 0, 1], [0, -1, 0]],
                  [[1, 0, -1], [1, -1, 1], [0, 1, 0]],
                  [[0, 0, 1], [-1, -1, -1], [-1, 0, 0]]],
                 [[[0, 1, 0], [1, -1, -1], [1, 1, -1]],
                  [[-1, 0, 1], [-1, -1, 1], [1, 1, 0]],
                  [[0, 1, 1], [-1, 1, -1], [-1, -1, 0]]]],
         target=[[[[1, -3, -1], [-4, 1, -6], [-3, -2, -1]],
                  [[-7, 1, 0], [3, -3, 2], [1, 3, -2]]]]),
]

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "conv2d_known_answer.json"), "w") as f:
        json.dump({"source": "transcribed from /root/reference/benchmarks/convolution/conv2d_common.nim:128-283",
                   "cases": CASES}, f, indent=1)
    print("wrote %d cases" % len(CASES))
