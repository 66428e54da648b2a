package PhyloProf::B200::Host;

# Host-side logic shared by the B200 command-line drop-ins (perl/bin/*_b200.pl): everything the reference's drivers do
# around the hot path -- reading BLAST caches / BLAST -m 8 tables / hmmsearch 3 reports, building depth-prefix query
# profiles, choosing p, ranking by printed score, formatting -- written from scratch against the observable behaviour of
#   bin/ppp.pl, bin/ppp2d.pl, bin/ppp_cross_score.pl, bin/ppp_hmmer.pl, bin/ppp_cutoff.pl,
#   lib/PhyloProf/BlastResult.pm, lib/PhyloProf/HMMERResult.pm, lib/SeqToolBox/HMMER/Parser.pm
# (cited per function).  The per-pair sweep itself is never done here: every scorer goes through PhyloProf::B200 (XS ->
# libppp_b200.so).  gi -> taxon lookups stay with the reference's own SeqToolBox::Taxonomy (its SQLite database), which
# the caller passes in.
use strict;
use warnings;
use Carp;
use Exporter 'import';
our @EXPORT_OK = qw(read_bla_list log10_3f handle_negative_zero parse_hmmsearch3 above_cutoff m8_rows m8_gi
    blast_target_list hmmer_target_list blast_prefix_queries hmmer_prefix_queries best_depth score_prefixes perl_num15);

# "%.15g": Perl's default stringification, which a value goes through whenever the reference writes it to a pipe or file
sub perl_num15 { return 0 + sprintf( "%.15g", $_[0] ) }

# sprintf("%.3f", log($score)/log(10)) -- dies on 0 as the reference does (bin/ppp.pl:476, bin/ppp_cross_score.pl:511)
sub log10_3f {
    my $score = shift;
    die "Can't take log of 0\n" if $score == 0;
    return sprintf( "%.3f", log($score) / log(10) );
}

# bin/ppp.pl:595-604: "-0.000" -> "0.000"
sub handle_negative_zero { my $s = shift; return $s =~ /^-([0.]+)$/ ? $1 : $s }

# One gene's cached BLAST hits -> the ordered taxon list the reference scores (bin/ppp.pl:606-673 + BlastResult.pm:192-237):
# subjects keep their first-occurrence order, a subject's LAST taxon wins, subjects with a false taxon are dropped,
# different subjects of one taxon stay as duplicates (they count in the depth).
sub read_bla_list {
    my $path = shift;
    open( my $fh, "<", $path ) or die "Can't open $path\n";
    my ( @order, %taxon );
    while ( my $line = <$fh> ) {
        chomp $line;
        my @f = split( /\t/, $line );
        my $subject = defined $f[1] ? $f[1] : '';
        push @order, $subject unless exists $taxon{$subject};
        $taxon{$subject} = $f[3];
    }
    close($fh);
    return [ grep { $_ } map { $taxon{$_} } @order ];
}

# ---------------------------------------------------------------------------------------------- hmmsearch 3 reports
# The per-sequence table of a hmmsearch 3 report (what lib/SeqToolBox/HMMER/Parser.pm:358-422 reads): rows between the
# column header that follows "Query:" and the "Domain annotation" / "Domain and alignment" section; the inclusion-threshold
# marker and blank lines inside the table are skipped.  Columns: E-value score bias | E-value score bias | exp N | name desc.
# Returns { full => {name => score}, domain => {name => best-domain score}, desc => {name => " words ..."} }.
sub parse_hmmsearch3 {
    my $file = shift;
    open( my $fh, "<", $file ) or die "Can't open $file\n";
    my ( %full, %domain, %desc );
    my $in = 0;
    while ( my $line = <$fh> ) {
        chomp $line;
        if ( $in && $line =~ /^Domain\s(?:and\salignment|annotation)/ ) { last }
        if ( !$in && $line =~ /^Query:/ ) {
            while ( my $skip = <$fh> ) { last if $skip =~ /^\s+E-value/ }
            my $dashes = <$fh>;
            $in = 1;
            next;
        }
        next unless $in;
        next if $line =~ /inclusion\sthreshold/ || $line eq "";
        ( my $l = $line ) =~ s/^\s+|\s+$//g;
        my @f = split( /\s+/, $l );
        my $name = $f[8];
        next unless defined $name;
        $full{$name}   = $f[1];
        $domain{$name} = $f[4];
        $desc{$name}   = join( "", map { " $_" } grep { $_ } @f[ 9 .. $#f ] );
    }
    close($fh);
    return { full => \%full, domain => \%domain, desc => \%desc };
}

# names with score > cutoff, best first (Parser.pm:322-356 get_above_cutoff_h3 / get_above_cutoff_domain_h3).  The reference
# orders equal scores by Perl's hash order; here they are ordered by name.
sub above_cutoff {
    my ( $scores, $cutoff ) = @_;
    $cutoff ||= 0;
    return [ grep { $scores->{$_} > $cutoff } sort { $scores->{$b} <=> $scores->{$a} || $a cmp $b } keys %$scores ];
}

# ---------------------------------------------------------------------------------------------- BLAST -m 8 tables
# non-comment, non-empty rows as 12-field lists (bin/ppp_cross_score.pl:361-370, 431-442)
sub m8_rows {
    my $file = shift;
    open( my $fh, "<", $file ) or die "Can't open $file\n";
    my @rows;
    while ( my $line = <$fh> ) {
        next if $line =~ /^\#/;
        chomp $line;
        next unless $line;
        my @f = split( /\t/, $line );
        die "Wrong BLAST result format. You should run BLAST with -m 8/9\n" if @f != 12;
        push @rows, \@f;
    }
    close($fh);
    return \@rows;
}

sub m8_gi {
    my $subject = shift;
    die "Could not parse gi from file\n" unless defined $subject && $subject =~ /gi\|(\d+)/ && $1;
    return $1;
}

# target list of a BLAST table (bin/ppp_cross_score.pl:350-413 + BlastResult.pm:192-237): one entry per row whose subject gi
# has a taxon -- repeated subjects are NOT merged (the reference's `exists $data{$f[11]}` tests the bit score as a key, so its
# de-duplication never fires) -- each holding the subject's LAST taxon.
sub blast_target_list {
    my ( $rows, $taxonomy ) = @_;
    my ( @order, %taxon );
    foreach my $f (@$rows) {
        my $t = $taxonomy->get_taxon( m8_gi( $f->[1] ) );
        next unless $t;
        push @order, $f->[1];
        $taxon{ $f->[1] } = $t;
    }
    return [ map { $taxon{$_} } @order ];
}

# target of a hmmsearch report (bin/ppp_cross_score.pl:326-348 / bin/ppp_hmmer.pl:246-296 + HMMERResult.pm:162-242): hits above
# the cut-off, best first; the profile is the HASH taxon => rank of its first hit, which the scorer walks in rank order
# (ppp.pm:147-149) -- i.e. the list of DISTINCT taxa in first-hit order (depth counts distinct taxa).
sub hmmer_target_list {
    my ( $hits, $taxonomy ) = @_;
    my ( @list, %seen );
    foreach my $name (@$hits) {
        my $t = $taxonomy->get_taxon($name);
        next unless $t;
        next if $seen{$t}++;
        push @list, $t;
    }
    return \@list;
}

# depth-prefix queries of a BLAST table (bin/ppp_cross_score.pl:415-546): at every row that brings a new taxon (subject to
# --skip / --start-depth / --end-depth) the query is {seen taxa => 1} and p = line_count / n_taxa * scale.
# Returns [ [line_count, gi, [taxa seen so far], p], ... ].
sub blast_prefix_queries {
    my ( $rows, $taxonomy, $n_taxa, %o ) = @_;
    my ( @out, @seen, %seen );
    my $line_count = 0;
    foreach my $f (@$rows) {
        ++$line_count;
        my $gi = m8_gi( $f->[1] );
        my $t  = $taxonomy->get_taxon($gi);
        next unless $t;
        next if $seen{$t}++;
        push @seen, $t;
        next if $line_count % ( $o{skip} || 1 );
        my $tc = scalar @seen;
        last if $o{end} && $tc > $o{end};
        next unless $tc >= ( defined $o{start} ? $o{start} : 1 );
        my $pr = $line_count / $n_taxa;
        $pr = $pr * $o{scale} if $o{scale};
        die "Error in calculating probability\n" unless $pr > 0;
        push @out, [ $line_count, $gi, [@seen], $pr ];
    }
    return \@out;
}

# the same over the hits of a hmmsearch report (bin/ppp_cross_score.pl:549-679); the gi of a depth is not recorded there
sub hmmer_prefix_queries {
    my ( $hits, $taxonomy, $n_taxa, %o ) = @_;
    my ( @out, @seen, %seen );
    my $line_count = 0;
    foreach my $name (@$hits) {
        ++$line_count;
        m8_gi($name);    # dies like the reference when the name carries no gi
        my $t = $taxonomy->get_taxon($name);
        next unless $t;
        next if $seen{$t}++;
        push @seen, $t;
        next if $line_count % ( $o{skip} || 1 );
        my $tc = scalar @seen;
        last if $o{end} && $tc > $o{end};
        next unless $tc >= ( defined $o{start} ? $o{start} : 1 );
        my $pr = $line_count / $n_taxa;
        $pr = $pr * $o{scale} if $o{scale};
        die "Error in calculating probability\n" unless $pr > 0;
        push @out, [ $line_count, undef, [@seen], $pr ];
    }
    return \@out;
}

# every depth-prefix query against ONE ordered target list in one GPU call (the reference writes a temporary profile file
# and runs the Perl sweep once per depth, bin/ppp_cross_score.pl:471-503, 681-699).  $collapse (optional) = the -l rank
# collapse of ppp.pm:328-370 applied to the target's taxa and to the query's (Profile.pm:455-466).
# Returns [ [score, yes, number, depth], ... ] per query.
sub score_prefixes {
    my ( $gpu, $queries, $target, $collapse ) = @_;
    return [] unless @$queries;
    my @qh;
    foreach my $q (@$queries) {
        my %h;
        foreach my $t ( @{ $q->[2] } ) {
            my $c = $collapse ? $collapse->($t) : $t;
            $h{$c} = 1 if $c;
        }
        push @qh, \%h;
    }
    my @tl = $collapse ? map { $collapse->($_) } @$target : @$target;
    die "Taxon not found" if grep { !$_ } @tl;    # ppp.pm:179
    my %u;
    foreach my $h (@qh) { $u{$_} = 1 for keys %$h }
    my @universe = sort keys %u;
    my $hits = $gpu->score_ranked( \@universe, \@qh, [ \@tl ], -probs => [ map { $_->[3] } @$queries ] );
    return [ map { [ @{$_}[ 2 .. 5 ] ] } @$hits ];
}

# the best depth of one direction: the first depth whose PRINTED %.3f log-score is strictly lower than the running best
# (bin/ppp_cross_score.pl:505-538).  Returns [line_count, gi, [log_score, yes, number, depth, "%.2f" p]] or undef.
sub best_depth {
    my ( $queries, $results ) = @_;
    my $best;
    for my $i ( 0 .. $#$queries ) {
        my ( $score, $yes, $number, $depth ) = @{ $results->[$i] };
        my @row = ( log10_3f($score), $yes, $number, $depth, sprintf( "%.2f", $queries->[$i][3] ) );
        $best = [ $queries->[$i][0], $queries->[$i][1], \@row ] if !$best || $best->[2][0] > $row[0];
    }
    return $best;
}

1;
