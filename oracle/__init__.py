"""CPU oracle for the Pix2Pose per-detection inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pix2pose_b200/`` (the product) may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs do, and only as the checker / the timed CPU arm.

Parity status: the reference has no tests, golden vectors or fixtures and
its modules cannot be imported here (keras 2.2.1 / tensorflow-gpu 1.9 / scikit-image at
module level).  What IS pinned by the reference's own code, executed in the build container
and committed as golden vectors (scripts under tests/golden/):

* ``pix2pose.get_boxes``, ``pix2pose.pnp_ransac`` and the whole ``pix2pose.est_pose`` control flow
  (recognition.py:28-224): the method definitions are compiled out of the reference file in
  memory and run with numpy / cv2; ``resize`` and ``generator_train.predict`` are injected
  (oracle resize, planted analytic maps).  tests/test_reference_golden.py: the oracle returns
  exactly the same boxes, crops, masks, R|t, inlier fractions and sentinels.
* ``pix2pose_util/common_util.py`` ``getXYZ`` / ``get_normal`` (imports here unmodified):
  tests/test_oracle_depth.py.
* the per-image loop of tools/5_evaluation_bop_basic.py:281-349 (those source lines exec'd on seeded
  detections): tests/test_evaluation_host.py checks pix2pose_b200/evaluation.py against it.
* the network TOPOLOGY: ``ae_model.aemodel_unet_resnet50`` / ``aemodel_unet_prob`` and
  ``resnet50_mod.ResNet50`` are executed on tests/fake_keras.py (a stand-in for the Keras functional
  API whose layers evaluate with net_oracle's primitives); net_oracle's hand-written forward gives
  the same bits as the graph the reference code builds, and the construction order of the
  weight-carrying layers is the one the Keras-HDF5 importer assumes (tests/test_oracle_net.py).

*Parity unpinned* remains for the two third-party semantics that cannot run here: the
Keras/TensorFlow per-layer arithmetic (TF 'same' padding, Conv2DTranspose cropping, BatchNorm epsilon, LeakyReLU
alpha as restated in net_oracle.py; algebraic self-checks only) and
scikit-image ``resize`` (resize_oracle.py, documented behaviour of skimage 0.14-0.18).  The PnP
itself is the real ``cv2.solvePnPRansac`` (container OpenCV 4.13.0; reference pins 3.4.2.17).
"""
