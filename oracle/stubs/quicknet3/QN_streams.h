/* intentionally empty: everything the hot path needs is in QuickNet.h (oracle stub) */
