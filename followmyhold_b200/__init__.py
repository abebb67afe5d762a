"""B200-native guidance + alignment hot path of FollowMyHold."""
__version__ = "0.1.0"
