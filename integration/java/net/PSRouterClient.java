package net;

/**
 * Drop-in note for net/PSRouterClient.java (:33 ctor, :60 getList, :93 updateList, :125 push,
 * :131 barrier).  On one B200 box the router's job — bucket keys by Router.shard(key), one batched
 * RPC per shard, merge — is done on the devices: ps_shard_route_dev buckets by
 * ps_owner_of(key, nGPU), NCCL all-to-all moves the buckets over NVLink, the owner runs
 * ps_model_shard_lookup_dev / ps_model_shard_apply_dev.  A JVM deployment runs one worker
 * process per GPU (the reference's "-Dmode=dist" worker, README.md:78-94) and calls the sequence
 * of INTEGRATION.md §4 through PsNative; no gRPC PServer process is needed on-box.
 */
public class PSRouterClient extends PSClient {
	public PSRouterClient(Router router) { super(); }
}
