package net;

import nativeps.PsNative;

/**
 * A net/Router.java:5 for the GPU-resident store: shard(key) is the owner the device-side route kernels compute for the same key
 * (ps_key_owner → ps_owner_of(pack(field, id), nShards), include/ps_spec.h) — the stock net/Mod.java hashes the key STRING and goes
 * negative for about 29 % of keys.  Keys every shard holds (wide and dense parameters are replicated) go to shard 0.
 * SOURCE ONLY: no JDK in the build image.
 */
public class NativeRouter implements Router {
	final int nShards;
	public NativeRouter(int nShards) { this.nShards = nShards; }
	public int shard(String key) {
		int o = PsNative.keyOwner(key, nShards);
		return o < 0 ? 0 : o;
	}
	/** true when every shard holds the key (an update must reach all of them) */
	public boolean replicated(String key) { return PsNative.keyOwner(key, nShards) < 0; }
}
