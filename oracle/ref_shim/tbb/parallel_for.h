// empty stand-in: the registration classes never use this header's contents (oracle build only)
