"""B200-native ORB front end (extraction + Hamming matching) behind the reference's API."""
