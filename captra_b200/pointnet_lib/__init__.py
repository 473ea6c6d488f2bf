"""Mirror of the reference's network/models/pointnet_lib package (only pointnet2_utils is live)."""
