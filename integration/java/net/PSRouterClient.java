package net;

/**
 * Drop-in for net/PSRouterClient.java (:33 ctor, :60 getList, :93 updateList, :125 push, :131 barrier).  The router's job —
 * bucket keys by Router.shard(key), one batched call per shard in parallel, merge (PSRouterClient.java:60-122) — is done on
 * the devices: every rank's route kernel buckets its batch's keys by ps_owner_of(key, nGPU) and stores each bucket straight
 * into the owner GPU's mailbox over NVLink; the owner's lookup kernel answers into the requester's mailbox (p2p.cu).  A JVM
 * deployment runs one worker process per GPU (the reference's "-Dmode=dist" worker) and calls ps_model_p2p_submit /
 * ps_model_collect through PsNative; host-side list access inherits PSClient's batched calls against the local shard.
 */
public class PSRouterClient extends PSClient {
	public PSRouterClient(Router router) { super(); }
	public PSRouterClient() { super(); }
}
