"""``TwentyFortyEight`` environment class — mirrors ``twentyfortyeight/envs`` of the reference."""
