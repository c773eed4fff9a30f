"""Drop-in shim: with this directory on PYTHONPATH in place of the reference's src/, ``from models.conv_tasnet import
ConvTasNet`` resolves to the sm_100a implementation (ctn_b200.models.*)."""
