"""ResnetBlock2D / TemporalResnetBlock are touched only by the UNet family (pinned through
oracle/unet.py); importing the module must succeed."""
